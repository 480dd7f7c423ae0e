"""ctypes binding of oracle/libtlag_cpu.so (ORACLE O2 / CPU baseline -- test infrastructure).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class CpuModel(C.Structure):
    _fields_ = [("W", C.c_uint32), ("code", C.c_void_p), ("code_len", C.c_uint32),
                ("entry_inv", C.c_uint32), ("entry_next", C.c_uint32),
                ("cpool", C.c_void_p), ("cpool_len", C.c_uint32),
                ("layout", C.c_void_p), ("n_slots", C.c_uint32),
                ("frame_words", C.c_uint32), ("unpacked_words", C.c_uint32), ("n_invariants", C.c_uint32),
                ("flags", C.c_uint32), ("table_log2", C.c_uint32), ("max_states", C.c_uint64)]


class CpuResult(C.Structure):
    _fields_ = [("verdict", C.c_int32), ("detail", C.c_int32), ("detail2", C.c_int32), ("pad", C.c_int32),
                ("state_idx", C.c_uint64), ("generated", C.c_uint64), ("distinct", C.c_uint64),
                ("depth", C.c_uint64), ("init_states", C.c_uint64), ("fp_xor", C.c_uint64), ("fp_sum", C.c_uint64),
                ("seconds", C.c_double), ("n_levels", C.c_uint64), ("level_sizes", C.c_uint64 * 4096),
                ("level_xor", C.c_uint64 * 4096), ("level_sum", C.c_uint64 * 4096),
                ("level_generated", C.c_uint64 * 4096)]


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        p = os.path.join(_HERE, "libtlag_cpu.so")
        if not os.path.exists(p):
            build()
        _LIB = C.CDLL(p)
        _LIB.tlagcpu_run.restype = C.c_int
        _LIB.tlagcpu_run2.restype = C.c_int
        _LIB.tlagcpu_probe_batch.restype = C.c_double
        _LIB.tlagcpu_fingerprint.restype = C.c_uint64
    return _LIB


def run(cm, init_words: np.ndarray, n_threads=1, deadlock=True, max_states=1 << 22, stop_after=0,
        want_states=False, exact=False, max_levels=0):
    """BFS of a CompiledModel on host cores.  Returns dict(verdict, generated, distinct, depth, ...).
    exact=True: one worker in FIFO order that stops at the first Assert failure / deadlock, i.e. the counts TLC's
    single worker prints at the moment of the error (`queue` = states discovered but not yet dequeued).
    max_levels=L: depth-bounded prefix -- levels 1..L-1 are expanded, the result holds every state of levels 1..L.
    `level_digests[k]` = (xor, sum, generated) cumulative after level k+1 exists: one run pins every shorter prefix."""
    L = lib()
    code = np.ascontiguousarray(cm.code, dtype=np.uint64)
    cpool = np.ascontiguousarray(cm.cpool, dtype=np.int32)
    layout = np.ascontiguousarray(cm.layout, dtype=np.int32)
    init = np.ascontiguousarray(init_words, dtype=np.uint32).reshape(-1, cm.W)
    m = CpuModel(cm.W, code.ctypes.data, len(code), cm.entries["inv"], cm.entries["next"],
                 cpool.ctypes.data, len(cpool), layout.ctypes.data, layout.shape[0],
                 cm.frame_words, cm.state_words_unpacked, len(cm.invariants),
                 (1 if deadlock else 0) | (4 if exact else 0), 0, max_states)
    res = CpuResult()
    states = None
    sp, cap = None, 0
    if want_states:
        states = np.zeros((max_states, cm.W), dtype=np.uint32)
        sp, cap = states.ctypes.data_as(C.c_void_p), max_states
    rc = L.tlagcpu_run2(C.byref(m), init.ctypes.data_as(C.c_void_p), C.c_uint64(init.shape[0]), C.c_int(n_threads),
                        C.c_uint64(stop_after), C.c_uint64(max_levels), C.byref(res), sp, C.c_uint64(cap))
    if rc != 0:
        raise RuntimeError(f"tlagcpu_run failed: {rc}")
    out = dict(verdict=res.verdict, detail=res.detail, detail2=res.detail2, state_idx=res.state_idx,
               generated=res.generated, distinct=res.distinct, depth=res.depth, init_states=res.init_states,
               fp_xor=res.fp_xor, fp_sum=res.fp_sum, seconds=res.seconds,
               levels=[int(res.level_sizes[i]) for i in range(res.n_levels)],
               level_digests=[(int(res.level_xor[i]), int(res.level_sum[i]), int(res.level_generated[i]))
                              for i in range(res.n_levels)])
    if exact and out["verdict"] in (2, 3):
        out["queue"] = out["distinct"] - out["state_idx"] - 1
    if want_states:
        out["states"] = states[:res.distinct].copy()
    return out


def digest(states: np.ndarray, W: int):
    s = np.ascontiguousarray(states, dtype=np.uint32)
    out = (C.c_uint64 * 2)()
    lib().tlagcpu_digest(s.ctypes.data_as(C.c_void_p), C.c_uint64(s.size // W), C.c_int(W), out)
    return int(out[0]), int(out[1])


def fingerprint(words: np.ndarray):
    w = np.ascontiguousarray(words, dtype=np.uint32)
    return int(lib().tlagcpu_fingerprint(w.ctypes.data_as(C.c_void_p), C.c_int(w.size)))


def probe_batch(states: np.ndarray, W: int, table_log2: int, n_threads: int):
    s = np.ascontiguousarray(states, dtype=np.uint32)
    n = s.size // W
    flags = np.zeros(n, dtype=np.uint8)
    dt = lib().tlagcpu_probe_batch(s.ctypes.data_as(C.c_void_p), C.c_uint64(n), C.c_int(W), C.c_uint(table_log2),
                                   C.c_int(n_threads), flags.ctypes.data_as(C.c_void_p))
    return flags, dt
