"""Serialisation of compiled models (.tlagz = numpy .npz): bytecode image, constant pool,
packed layout, encoded initial states and the typing information needed to decode states.

Used for (a) the committed parity fixtures under tests/golden/ (the .tla sources of the
reference corpus do not travel to the GPU box, their compiled form does) and (b) caching."""
from __future__ import annotations

import io
import json
import pickle

import numpy as np

from .compile.lower import CompiledModel


def save_compiled(path, cm: CompiledModel, init_words: np.ndarray, expected: dict | None = None, info: dict | None = None):
    meta = dict(W=cm.W, frame_words=cm.frame_words, entries=cm.entries,
                state_words_unpacked=cm.state_words_unpacked, vars=cm.vars, invariants=cm.invariants,
                actions=[[a[0], list(a[1]), a[2]] for a in cm.actions],
                asserts=[[a[0], list(a[1])] for a in cm.asserts],
                module_name=getattr(cm, "module_name", ""), state_bits=getattr(cm, "state_bits", 0),
                segments=getattr(cm, "segments", None), blocks=getattr(cm, "blocks", None),
                expected=expected or {}, info=info or {})
    typing = pickle.dumps(dict(var_types=cm.var_types, var_off=cm.var_off, atoms=cm.atoms.vals))
    with open(path, "wb") as f:
        np.savez_compressed(f, code=np.asarray(cm.code, dtype=np.uint64), cpool=np.asarray(cm.cpool, dtype=np.int32),
                            layout=np.asarray(cm.layout, dtype=np.int32),
                            init=np.asarray(init_words, dtype=np.uint32),
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
                            typing=np.frombuffer(typing, dtype=np.uint8))


def load_compiled(path):
    """-> (CompiledModel, init_words, expected dict, info dict)"""
    from .compile.types import Atoms, Codec
    z = np.load(path, allow_pickle=False)
    meta = json.loads(bytes(z["meta"]).decode())
    cm = CompiledModel()
    cm.code, cm.cpool, cm.layout = z["code"], z["cpool"], z["layout"]
    cm.W, cm.frame_words = meta["W"], meta["frame_words"]
    cm.entries = meta["entries"]
    cm.state_words_unpacked = meta["state_words_unpacked"]
    cm.vars, cm.invariants = meta["vars"], meta["invariants"]
    cm.actions = [(a[0], tuple(a[1]), a[2]) for a in meta["actions"]]
    cm.asserts = [(a[0], tuple(a[1])) for a in meta["asserts"]]
    cm.module_name = meta.get("module_name", "")
    cm.state_bits = meta.get("state_bits", 0)
    cm.segments = meta.get("segments")
    cm.blocks = [tuple(b) for b in meta["blocks"]] if meta.get("blocks") else None
    ty = pickle.loads(bytes(z["typing"]))
    cm.var_types, cm.var_off = ty["var_types"], ty["var_off"]
    at = Atoms()
    for v in ty["atoms"][1:]:
        at.id(v)
    cm.atoms = at
    cm.codec = Codec(at)
    cm.n_off, cm.p_off = 0, cm.state_words_unpacked
    return cm, z["init"], meta.get("expected", {}), meta.get("info", {})
