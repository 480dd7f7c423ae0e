"""ctypes binding of the C-ABI engine library (include/tlag.h, csrc/libtlag.so).

The product path has no CPU fallback: if the CUDA extension is missing or no GPU is
visible, construction fails loudly (EngineUnavailable)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TLAG_LIB") or os.path.join(_HERE, "csrc", "libtlag.so")   # TLAG_LIB: tuning builds
_LIB = None

F_DEADLOCK_CHECK = 1
F_KEEP_GOING = 2
F_EXACT = 4
V_OK, V_INVARIANT, V_ASSERT, V_DEADLOCK, V_EVAL_ERROR, V_RUNNING = 0, 1, 2, 3, 4, 5


class EngineUnavailable(RuntimeError):
    pass


class EngineError(RuntimeError):
    pass


class TlagModel(C.Structure):
    _fields_ = [("words_per_state", C.c_uint32), ("code", C.c_void_p), ("code_len", C.c_uint32),
                ("entry_inv", C.c_uint32), ("entry_next", C.c_uint32),
                ("cpool", C.c_void_p), ("cpool_len", C.c_uint32),
                ("layout", C.c_void_p), ("n_slots", C.c_uint32),
                ("frame_words", C.c_uint32), ("unpacked_words", C.c_uint32),
                ("n_invariants", C.c_uint32), ("n_actions", C.c_uint32),
                ("table_slots_log2", C.c_uint32), ("max_states", C.c_uint64),
                ("flags", C.c_uint32), ("device", C.c_int32)]


class WaveStats(C.Structure):
    _fields_ = [("level", C.c_uint64), ("expanded", C.c_uint64), ("generated", C.c_uint64),
                ("discovered", C.c_uint64), ("distinct_total", C.c_uint64), ("generated_total", C.c_uint64),
                ("kernel_ms", C.c_float), ("verdict", C.c_int32)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class TlagResult(C.Structure):
    _fields_ = [("verdict", C.c_int32), ("detail", C.c_int32), ("detail2", C.c_int32), ("reserved", C.c_int32),
                ("state_idx", C.c_uint64), ("generated", C.c_uint64), ("distinct", C.c_uint64),
                ("queue_left", C.c_uint64), ("depth", C.c_uint64), ("init_states", C.c_uint64),
                ("fp_collision_estimate", C.c_double), ("device_seconds", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


EXPORTS = ["tlag_create", "tlag_seed", "tlag_step", "tlag_run", "tlag_result_now", "tlag_trace",
           "tlag_read_states", "tlag_digest", "tlag_probe_batch", "tlag_probe_batch_device", "tlag_reset_table", "tlag_restart",
           "tlag_kernel_launches", "tlag_destroy", "tlag_last_error", "tlag_version",
           "tlag_frontier", "tlag_expand_route", "tlag_insert_records", "tlag_advance_level",
           "tlag_p2p_init", "tlag_p2p_attach", "tlag_p2p_level", "tlag_p2p_rollback", "tlag_read_link", "tlag_set_rank", "tlag_set_owner_words"]


def build_library(verbose=False):
    """nvcc cross-compiles for sm_100a without a GPU (driver `build()` check)."""
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc")] + ([] if verbose else ["-s"]))


def _bind(path):
    L = C.CDLL(path)
    L.tlag_last_error.restype = C.c_char_p
    L.tlag_version.restype = C.c_char_p
    L.tlag_kernel_launches.restype = C.c_uint64
    L.tlag_destroy.restype = None
    for fn in EXPORTS:
        getattr(L, fn)
    return L


def load_library():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise EngineUnavailable(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                    f"(the product has no CPU fallback)")
        _LIB = _bind(LIB_PATH)
    return _LIB


FRAME_CLASSES = (64, 128, 256, 512, 1024, 2048, 4096, 8192)     # csrc/tlag_engine.cu: kFrameClasses
NATIVE_DIR = os.path.join(_HERE, "csrc", "native")
_NATIVE_LIBS = {}


def _sliced_tag(cm, scalar):
    import hashlib
    from .compile.sliced import model_key
    h = hashlib.sha256()
    for rel in ("csrc/tlag_engine.cu", "csrc/tlag_dev.cuh", "csrc/tlag_dev2.cuh", "csrc/tlag_vm.h", "csrc/tlag_vm_exec.inc",
                "compile/sliced.py", "../include/tlag.h"):
        with open(os.path.join(_HERE, rel), "rb") as f:
            h.update(f.read())
    h.update(repr((getattr(cm, "segments", None), os.environ.get("TLAG_SL_OCC", ""), os.environ.get("TLAG_SL_BLOCK", ""),
                   os.environ.get("TLAG_SL_MIN_SLICE", ""))).encode())
    if os.environ.get("TLAG_CSRC_DIR"):                 # experiments built from a modified copy of csrc/
        h.update(os.environ["TLAG_CSRC_DIR"].encode())
    return f"sl_{model_key(cm)}_{h.hexdigest()[:8]}{'_s' if scalar else ''}"


def sliced_form(cm):
    """scalar form (frame words as C locals) when the model allows it, else the array form; TLAG_SL_FORM overrides."""
    from .compile.sliced import Emitter, SliceError
    want = os.environ.get("TLAG_SL_FORM", "auto")
    if want == "array":
        return False
    try:
        Emitter(cm, scalar=True)
        return True
    except SliceError:
        if want == "scalar":
            raise
        return False


def sliced_library_path(cm, scalar=None):
    if scalar is None:
        scalar = sliced_form(cm)
    return os.path.join(NATIVE_DIR, f"libtlag_{_sliced_tag(cm, scalar)}.so")


def build_sliced_library(cm, force=False, verbose=False, scalar=None, jobs=None):
    """Model-specialised engine library, sliced form (compile/sliced.py): one kernel per invariant and per disjunct of
    Next, same engine source and C ABI.  The slices are spread over several translation units that nvcc compiles in
    parallel (-dc) next to the engine's own unit; built in-tree (csrc/native/) so that it travels with the snapshot."""
    from concurrent.futures import ThreadPoolExecutor
    from .compile.sliced import emit_parts
    if scalar is None:
        scalar = sliced_form(cm)
    os.makedirs(NATIVE_DIR, exist_ok=True)
    tag = _sliced_tag(cm, scalar)
    so = os.path.join(NATIVE_DIR, f"libtlag_{tag}.so")
    if os.path.exists(so) and not force:
        return so
    jobs = jobs or int(os.environ.get("TLAG_BUILD_JOBS", "0")) or min(8, os.cpu_count() or 1)
    work = os.path.join(NATIVE_DIR, tag)
    os.makedirs(work, exist_ok=True)
    defs_path = os.path.join(work, "defs.h")
    nparts = max(1, min(jobs, 1 + len(cm.code) // 1500))
    ms = os.environ.get("TLAG_SL_MIN_SLICE")
    defs, parts = emit_parts(cm, scalar, nparts, defs_path, min_slice=int(ms) if ms else None)
    with open(defs_path, "w") as f:
        f.write(defs)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    csrc = os.environ.get("TLAG_CSRC_DIR") or os.path.join(_HERE, "csrc")      # (experiments: a modified copy of csrc/)
    base = [nvcc, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
            "-I", csrc, "-dc"]
    if not scalar and cm.frame_words > 512:
        base.append("-DTLAG_SL_SEG_NOINLINE")
    for k in ("TLAG_SL_OCC", "TLAG_SL_BLOCK"):
        if os.environ.get(k):
            base.append(f"-D{k}={int(os.environ[k])}")
    # subroutines are compiled apart from the kernels that call them (relocatable device code): give them the register
    # budget of the kernels' __launch_bounds__ (tlag_dev.cuh: 256 threads x 4 CTAs per SM by default)
    regcap = 65536 // (int(os.environ.get("TLAG_SL_BLOCK", "256")) * int(os.environ.get("TLAG_SL_OCC", "4")))
    base.append(f"-maxrregcount={min(255, regcap)}")
    units = []
    for k, text in enumerate(parts):
        src = os.path.join(work, f"part{k}.cu")
        with open(src, "w") as f:
            f.write(text)
        units.append((base + ["-o", src[:-3] + ".o", src], src[:-3] + ".o"))
    units.append((base + [f'-DTLAG_SLICED_INC="{defs_path}"', "-o", os.path.join(work, "engine.o"),
                          os.path.join(csrc, "tlag_engine.cu")], os.path.join(work, "engine.o")))

    def run(cmd):
        p = subprocess.run(cmd, capture_output=not verbose, text=True)
        if p.returncode != 0:
            raise EngineError(f"nvcc failed for the sliced build of model {tag}: {(p.stderr or '')[-3000:]}")
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        list(ex.map(run, [u[0] for u in units]))
    run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so + ".tmp"] + [u[1] for u in units])
    os.replace(so + ".tmp", so)
    if os.environ.get("TLAG_KEEP_BUILD") != "1":       # generated sources + objects (regenerated from the model at will)
        import shutil
        shutil.rmtree(work, ignore_errors=True)
    return so


def load_sliced_library(cm, build=True):
    if os.environ.get("TLAG_NO_BUILD") == "1":      # GPU sessions: never spend box time in nvcc
        build = False
    so = sliced_library_path(cm)
    if so not in _NATIVE_LIBS:
        if not os.path.exists(so):
            if not build:
                raise EngineUnavailable(f"{so} is missing (sliced native build of this model)")
            build_sliced_library(cm)
        _NATIVE_LIBS[so] = _bind(so)
    return _NATIVE_LIBS[so]


class Engine:
    """One BFS engine instance on one GPU (mirrors tlag_engine)."""

    def __init__(self, cm, deadlock=True, device=0, table_log2=0, max_states=0, keep_going=False, native=None, exact=False):
        """native="sliced" (or True): the model-specialised sliced build -- one kernel per invariant / disjunct of Next,
        compile/sliced.py; built on demand with nvcc and cached in csrc/native/; False: the bytecode interpreter kernel;
        None: follow the environment variable TLAG_NATIVE (sliced / 0)."""
        if native is None:
            native = "sliced" if os.environ.get("TLAG_NATIVE", "0") in ("sliced", "1") else False
        if exact:               # TLAG_F_EXACT: TLC's single worker replayed on the device (interpreter kernel, one warp)
            native = False
        self.native = "sliced" if native else False
        self.L = load_sliced_library(cm) if self.native else load_library()
        self.cm = cm
        self._code = np.ascontiguousarray(cm.code, dtype=np.uint64)
        self._cpool = np.ascontiguousarray(cm.cpool, dtype=np.int32)
        self._layout = np.ascontiguousarray(cm.layout, dtype=np.int32)
        flags = (F_DEADLOCK_CHECK if deadlock else 0) | (F_KEEP_GOING if keep_going else 0) | (F_EXACT if exact else 0)
        m = TlagModel(cm.W, self._code.ctypes.data, len(self._code), cm.entries["inv"], cm.entries["next"],
                      self._cpool.ctypes.data, len(self._cpool), self._layout.ctypes.data, self._layout.shape[0],
                      cm.frame_words, cm.state_words_unpacked, len(cm.invariants), len(cm.actions),
                      table_log2, max_states, flags, device)
        self.h = C.c_void_p()
        rc = self.L.tlag_create(C.byref(m), C.byref(self.h))
        if rc != 0:
            msg = self.L.tlag_last_error(self.h).decode() if self.h else "tlag_create failed"
            if self.h:
                self.L.tlag_destroy(self.h)
                self.h = None
            if rc == -3:
                raise EngineUnavailable(f"CUDA engine unavailable: {msg}")
            raise EngineError(f"tlag_create: {msg}")

    def _ck(self, rc, what):
        if rc != 0:
            raise EngineError(f"{what}: rc={rc}: {self.L.tlag_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.tlag_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def seed(self, init_words: np.ndarray):
        a = np.ascontiguousarray(init_words, dtype=np.uint32).reshape(-1, self.cm.W)
        self._ck(self.L.tlag_seed(self.h, a.ctypes.data_as(C.c_void_p), C.c_uint64(a.shape[0])), "tlag_seed")

    def step(self) -> dict:
        ws = WaveStats()
        self._ck(self.L.tlag_step(self.h, C.byref(ws)), "tlag_step")
        return ws.as_dict()

    def run(self) -> dict:
        r = TlagResult()
        self._ck(self.L.tlag_run(self.h, C.byref(r)), "tlag_run")
        return r.as_dict()

    def result(self) -> dict:
        r = TlagResult()
        self._ck(self.L.tlag_result_now(self.h, C.byref(r)), "tlag_result_now")
        return r.as_dict()

    def trace(self, state_idx: int, cap=4096):
        W = self.cm.W
        states = np.zeros((cap, W), dtype=np.uint32)
        acts = np.zeros(cap, dtype=np.int32)
        n = C.c_uint32(cap)
        self._ck(self.L.tlag_trace(self.h, C.c_uint64(state_idx), states.ctypes.data_as(C.c_void_p),
                                   acts.ctypes.data_as(C.c_void_p), C.byref(n)), "tlag_trace")
        return states[:n.value].copy(), acts[:n.value].copy()

    def read_states(self, first: int, n: int) -> np.ndarray:
        out = np.zeros((n, self.cm.W), dtype=np.uint32)
        self._ck(self.L.tlag_read_states(self.h, C.c_uint64(first), C.c_uint64(n), out.ctypes.data_as(C.c_void_p)),
                 "tlag_read_states")
        return out

    def read_link(self, idx: int):
        """-> (state words, parent index or -1 for an initial state, action id, rank holding the parent)"""
        st = np.zeros(self.cm.W, dtype=np.uint32)
        par, meta = C.c_uint32(), C.c_uint32()
        self._ck(self.L.tlag_read_link(self.h, C.c_uint64(idx), st.ctypes.data_as(C.c_void_p), C.byref(par), C.byref(meta)),
                 "tlag_read_link")
        root = par.value == 0xFFFFFFFF
        return st, (-1 if root else par.value), (-1 if root else meta.value >> 8), meta.value & 0xFF

    def digest(self):
        x, s_ = C.c_uint64(), C.c_uint64()
        self._ck(self.L.tlag_digest(self.h, C.byref(x), C.byref(s_)), "tlag_digest")
        return int(x.value), int(s_.value)

    def probe_batch(self, states: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(states, dtype=np.uint32).reshape(-1, self.cm.W)
        flags = np.zeros(a.shape[0], dtype=np.uint8)
        self._ck(self.L.tlag_probe_batch(self.h, a.ctypes.data_as(C.c_void_p), C.c_uint64(a.shape[0]),
                                         flags.ctypes.data_as(C.c_void_p)), "tlag_probe_batch")
        return flags

    def probe_batch_device(self, d_states_ptr: int, n: int, d_flags_ptr: int) -> float:
        ms = C.c_float()
        self._ck(self.L.tlag_probe_batch_device(self.h, C.c_uint64(d_states_ptr), C.c_uint64(n),
                                                C.c_uint64(d_flags_ptr), C.byref(ms)), "tlag_probe_batch_device")
        return ms.value

    def reset_table(self):
        self._ck(self.L.tlag_reset_table(self.h), "tlag_reset_table")

    def restart(self):
        self._ck(self.L.tlag_restart(self.h), "tlag_restart")

    def launches(self) -> int:
        return int(self.L.tlag_kernel_launches(self.h))

    # multi-GPU building blocks
    def frontier(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self.L.tlag_frontier(self.h, C.byref(a), C.byref(b)), "tlag_frontier")
        return int(a.value), int(b.value)

    def set_owner_words(self, k):
        self._ck(self.L.tlag_set_owner_words(self.h, C.c_uint32(k)), "tlag_set_owner_words")

    def set_rank(self, n_ranks, rank):
        self._ck(self.L.tlag_set_rank(self.h, C.c_uint32(n_ranks), C.c_uint32(rank)), "tlag_set_rank")

    def expand_route(self, n_ranks, first, count, d_send_ptr, cap_records):
        counts = (C.c_uint64 * n_ranks)()
        ws = WaveStats()
        self._ck(self.L.tlag_expand_route(self.h, C.c_uint32(n_ranks), C.c_uint64(first), C.c_uint64(count),
                                          C.c_uint64(d_send_ptr), C.c_uint64(cap_records), counts, C.byref(ws)),
                 "tlag_expand_route")
        return [int(c) for c in counts], ws.as_dict()

    def insert_records(self, d_recv_ptr, n_records) -> int:
        nn = C.c_uint64()
        self._ck(self.L.tlag_insert_records(self.h, C.c_uint64(d_recv_ptr), C.c_uint64(n_records), C.c_uint32(0),
                                            C.byref(nn)), "tlag_insert_records")
        return int(nn.value)

    # peer-memory exchange (include/tlag.h: tlag_p2p_*)
    def p2p_init(self, n_ranks, rank, cap_records) -> bytes:
        h = (C.c_uint8 * 64)()
        self._ck(self.L.tlag_p2p_init(self.h, C.c_uint32(n_ranks), C.c_uint32(rank), C.c_uint64(cap_records), h), "tlag_p2p_init")
        return bytes(h)

    def p2p_attach(self, peer, handle: bytes):
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        self._ck(self.L.tlag_p2p_attach(self.h, C.c_uint32(peer), h), "tlag_p2p_attach")

    def p2p_level(self, n_chunks, chunk_states, expect_inbound=0):
        """-> wave stats dict, or None when one of this rank's send regions overflowed (see p2p_rollback)"""
        ws = WaveStats()
        rc = self.L.tlag_p2p_level(self.h, C.c_uint64(n_chunks), C.c_uint64(chunk_states), C.c_uint64(expect_inbound),
                                   C.byref(ws))
        if rc == -5:            # TLAG_EOVERFLOW
            return None
        self._ck(rc, "tlag_p2p_level")
        return ws.as_dict()

    def p2p_rollback(self):
        self._ck(self.L.tlag_p2p_rollback(self.h), "tlag_p2p_rollback")

    def advance_level(self) -> dict:
        ws = WaveStats()
        self._ck(self.L.tlag_advance_level(self.h, C.byref(ws)), "tlag_advance_level")
        return ws.as_dict()


class ProbeOnlyModel:
    """Minimal stand-in CompiledModel for K1-only use (tlag_probe_batch on W-word states)."""

    def __init__(self, W):
        self.W = W
        self.code = np.zeros(2, dtype=np.uint64)  # HALT, HALT
        self.cpool = np.zeros(1, dtype=np.int32)
        self.layout = np.array([[i, 32, 0] for i in range(W)], dtype=np.int32)
        self.entries = {"inv": 0, "next": 1}
        self.frame_words = 2 * W + 4
        self.state_words_unpacked = W
        self.invariants = []
        self.actions = []
