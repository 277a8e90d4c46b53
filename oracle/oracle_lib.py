"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the two CPU checkers.

  Port : oracle/_build/libmg_oracle.so      (oracle/mg_oracle.c, our restatement)
  Ref  : oracle/_ref/libmegalania_ref.so    (the UNMODIFIED reference sources + ref_harness.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package megalania_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libmg_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmegalania_ref.so")
REFERENCE_ROOT = "/root/reference"

# LZMAPacket, reference src/lzma_packet.h:13-17 (12 bytes, natural alignment)
PACKET_DTYPE = np.dtype({"names": ["type", "dist", "len"], "formats": ["u1", "<u4", "<u2"],
                         "offsets": [0, 4, 8], "itemsize": 12})
TRACE_DTYPE = np.dtype([("cost", "<u8"), ("flags", "<u4"), ("undo_count", "<u4")])
NUM_PROBS = 2615
MODEL_DTYPE = np.dtype({"names": ["probs", "ctx_state", "dists", "position", "cost"],
                        "formats": [("<u2", NUM_PROBS), "u1", ("<u4", 4), "<u8", "<u8"],
                        "offsets": [0, 5230, 5232, 5248, 5256], "itemsize": 5264})

LITERAL, MATCH, SHORT_REP, LONG_REP = 1, 2, 3, 4


def build(force: bool = False) -> None:
    """Compile the port, and the reference harness when /root/reference is present."""
    targets = []
    if force or not os.path.exists(PORT_SO):
        targets.append("port")
    if os.path.isdir(REFERENCE_ROOT) and (force or not os.path.exists(REF_SO)):
        targets += ["ref", "ref_cli"]
    if targets:
        subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def literal_slab(n: int) -> np.ndarray:
    slab = np.zeros(n, dtype=PACKET_DTYPE)
    slab["type"] = LITERAL
    slab["len"] = 1
    return slab


def _u8(data) -> np.ndarray:
    return np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class _Lib:
    prefix = ""

    def __init__(self, path: str):
        self.path = path
        self.lib = C.CDLL(path)
        p = self.prefix
        L = self.lib
        vp, sz, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
        getattr(L, p + "slab_cost").restype = u64
        getattr(L, p + "slab_cost").argtypes = [vp, sz, vp]
        getattr(L, p + "prefix_cost").restype = u64
        getattr(L, p + "prefix_cost").argtypes = [vp, sz, vp, sz]
        getattr(L, p + "model_after_prefix").restype = None
        getattr(L, p + "model_after_prefix").argtypes = [vp, sz, vp, sz, vp]
        getattr(L, p + "encode_slab").restype = sz
        getattr(L, p + "encode_slab").argtypes = [vp, sz, vp, vp, sz]
        getattr(L, p + "topk_many").restype = i32
        getattr(L, p + "topk_many").argtypes = [vp, sz, vp, i32, vp, sz, i32, vp, vp]
        getattr(L, p + "substring_count").restype = sz
        getattr(L, p + "substring_count").argtypes = [vp, sz, sz, sz]
        getattr(L, p + "heap_topk").restype = i32
        getattr(L, p + "heap_topk").argtypes = [vp, i32, i32, vp]

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def slab_cost(self, data, slab) -> int:
        d = _u8(data)
        return int(self._f("slab_cost")(_ptr(d), d.size, _ptr(slab)))

    def prefix_cost(self, data, slab, stop) -> int:
        d = _u8(data)
        return int(self._f("prefix_cost")(_ptr(d), d.size, _ptr(slab), stop))

    def model_after_prefix(self, data, slab, stop) -> np.ndarray:
        d = _u8(data)
        out = np.zeros(1, dtype=MODEL_DTYPE)
        self._f("model_after_prefix")(_ptr(d), d.size, _ptr(slab), stop, _ptr(out))
        return out[0]

    def encode_slab(self, data, slab) -> bytes:
        d = _u8(data)
        cap = d.size * 2 + 64
        out = np.zeros(cap, dtype=np.uint8)
        got = int(self._f("encode_slab")(_ptr(d), d.size, _ptr(slab), _ptr(out), cap))
        assert got <= cap
        return out[:got].tobytes()

    def topk_many(self, data, slab, state_mode, positions, k=20):
        d = _u8(data)
        pos = np.ascontiguousarray(positions, dtype=np.uint64)
        pops = np.zeros((pos.size, k), dtype=PACKET_DTYPE)
        counts = np.zeros(pos.size, dtype=np.int32)
        rc = self._f("topk_many")(_ptr(d), d.size, _ptr(slab), state_mode, _ptr(pos), pos.size, k,
                                  _ptr(pops), _ptr(counts))
        if rc != 0:
            raise ValueError("positions must be ascending live packet boundaries")
        return pops, counts

    def substring_count(self, data, pos, max_len=273) -> int:
        d = _u8(data)
        return int(self._f("substring_count")(_ptr(d), d.size, pos, max_len))

    def heap_topk(self, keys, k):
        keys = np.ascontiguousarray(keys, dtype=np.int32)
        out = np.zeros(k, dtype=np.int32)
        got = self._f("heap_topk")(_ptr(keys), keys.size, k, _ptr(out))
        return out[:got]


class Port(_Lib):
    prefix = "mgo_"

    def __init__(self):
        build()
        super().__init__(PORT_SO)
        L = self.lib
        vp, sz, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
        L.mgo_price_table.restype = C.POINTER(C.c_uint32)
        L.mgo_topk_many_priced.restype = i32
        L.mgo_topk_many_priced.argtypes = [vp, sz, vp, i32, vp, sz, i32, vp, vp, vp]
        L.mgo_srand.argtypes = [C.c_uint]
        L.mgo_rand.restype = i32
        L.mgo_chain_seed.restype = u64
        L.mgo_chain_seed.argtypes = [u64, u64]
        L.mgo_chain_rand31.restype = C.c_uint32
        L.mgo_chain_rand31.argtypes = [C.POINTER(u64)]
        L.mgo_anneal_epoch.restype = C.c_long
        L.mgo_anneal_epoch.argtypes = [vp, sz, vp, vp, C.POINTER(u64), i32, i32, C.c_uint,
                                       C.POINTER(u64), C.c_uint, i32, i32, i32, C.c_long,
                                       C.POINTER(u64), vp, C.c_long]
        L.mgo_slab_events.restype = sz
        L.mgo_slab_events.argtypes = [vp, sz, vp, vp, sz]
        L.mgo_greedy_slab.restype = None
        L.mgo_greedy_slab.argtypes = [vp, sz, vp]
        L.mgo_live_count.restype = sz
        L.mgo_live_count.argtypes = [vp, sz]
        L.mgo_slab_valid.restype = i32
        L.mgo_slab_valid.argtypes = [vp, sz, vp]
        L.mgo_set_finder_limits.restype = None
        L.mgo_set_finder_limits.argtypes = [sz, C.c_uint32]
        L.mgo_set_lc.restype = None
        L.mgo_set_lc.argtypes = [C.c_uint]

    def set_lc(self, lc: int = 0) -> None:
        """Literal context bits (process-global; 0 = the reference)."""
        self.lib.mgo_set_lc(lc)

    def set_finder_limits(self, window: int = 0, max_occ: int = 0) -> None:
        """Process-global: 0, 0 restores the reference's semantics."""
        self.lib.mgo_set_finder_limits(window, max_occ)

    def price_table(self) -> np.ndarray:
        p = self.lib.mgo_price_table()
        return np.ctypeslib.as_array(p, shape=(2048,)).copy()

    def topk_many_priced(self, data, slab, state_mode, positions, k=20):
        d = _u8(data)
        pos = np.ascontiguousarray(positions, dtype=np.uint64)
        pops = np.zeros((pos.size, k), dtype=PACKET_DTYPE)
        prices = np.zeros((pos.size, k), dtype=np.uint32)
        counts = np.zeros(pos.size, dtype=np.int32)
        rc = self.lib.mgo_topk_many_priced(_ptr(d), d.size, _ptr(slab), state_mode, _ptr(pos), pos.size,
                                           k, _ptr(pops), _ptr(prices), _ptr(counts))
        if rc != 0:
            raise ValueError("positions must be ascending live packet boundaries")
        return pops, prices, counts

    def rand_stream(self, seed, count):
        self.lib.mgo_srand(seed)
        return np.array([self.lib.mgo_rand() for _ in range(count)], dtype=np.int32)

    def chain_seed(self, seed, chain) -> int:
        return int(self.lib.mgo_chain_seed(seed, chain))

    def anneal_epoch(self, data, slab, best, best_cost, cur_cost, *, rng_mode=0, reseed=1,
                     seed=1673551, rng_state=0, step=0, num_iters=None, first_eval=0, evals=100,
                     max_attempts=None, trace_cap=None):
        """Returns (attempts, best_cost, cur_cost, rng_state, trace); slab/best edited in place."""
        d = _u8(data)
        if num_iters is None:
            num_iters = d.size
        if max_attempts is None:
            max_attempts = evals * 64 + 1024
        if trace_cap is None:
            trace_cap = max_attempts
        trace = np.zeros(trace_cap, dtype=TRACE_DTYPE)
        bc, cc, rs = C.c_uint64(best_cost), C.c_uint64(cur_cost), C.c_uint64(rng_state)
        attempts = self.lib.mgo_anneal_epoch(_ptr(d), d.size, _ptr(slab), _ptr(best), C.byref(bc),
                                             rng_mode, reseed, seed, C.byref(rs), step, num_iters,
                                             first_eval, evals, max_attempts, C.byref(cc), _ptr(trace),
                                             trace_cap)
        return int(attempts), int(bc.value), int(cc.value), int(rs.value), trace[:min(attempts, trace_cap)]

    def slab_events(self, data, slab) -> np.ndarray:
        d = _u8(data)
        cap = d.size * 12 + 64
        out = np.zeros(cap, dtype=np.uint16)
        got = int(self.lib.mgo_slab_events(_ptr(d), d.size, _ptr(slab), _ptr(out), cap))
        assert got <= cap
        return out[:got]

    def greedy_slab(self, data) -> np.ndarray:
        d = _u8(data)
        slab = literal_slab(d.size)
        self.lib.mgo_greedy_slab(_ptr(d), d.size, _ptr(slab))
        return slab

    def live_count(self, slab) -> int:
        return int(self.lib.mgo_live_count(_ptr(slab), slab.size))

    def slab_valid(self, data, slab) -> bool:
        d = _u8(data)
        return bool(self.lib.mgo_slab_valid(_ptr(d), d.size, _ptr(slab)))


class Ref(_Lib):
    """The unmodified reference, when oracle/_ref was built (here, or shipped to the GPU box)."""
    prefix = "mgref_"

    def __init__(self):
        build()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        super().__init__(REF_SO)
        L = self.lib
        vp, sz, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
        L.mgref_price_table.restype = C.POINTER(C.c_uint64)
        L.mgref_rand_stream.argtypes = [C.c_uint, i32, vp]
        L.mgref_anneal_epoch.restype = C.c_long
        L.mgref_anneal_epoch.argtypes = [vp, sz, vp, vp, C.POINTER(u64), i32, C.c_uint, C.c_uint, i32, i32,
                                         C.c_long, C.POINTER(u64), vp, C.c_long]
        L.mgref_sizeof_packet.restype = sz
        L.mgref_sizeof_state.restype = sz

    def price_table(self) -> np.ndarray:
        p = self.lib.mgref_price_table()
        return np.ctypeslib.as_array(p, shape=(2048,)).copy()

    def rand_stream(self, seed, count):
        out = np.zeros(count, dtype=np.int32)
        self.lib.mgref_rand_stream(seed, count, _ptr(out))
        return out

    def anneal_epoch(self, data, slab, best, best_cost, cur_cost, *, reseed=1, seed=1673551, step=0,
                     num_iters=None, evals=100, max_attempts=None, trace_cap=None):
        d = _u8(data)
        if num_iters is None:
            num_iters = d.size
        if max_attempts is None:
            max_attempts = evals * 64 + 1024
        if trace_cap is None:
            trace_cap = max_attempts
        trace = np.zeros(trace_cap, dtype=TRACE_DTYPE)
        bc, cc = C.c_uint64(best_cost), C.c_uint64(cur_cost)
        attempts = self.lib.mgref_anneal_epoch(_ptr(d), d.size, _ptr(slab), _ptr(best), C.byref(bc), reseed,
                                               seed, step, num_iters, evals, max_attempts, C.byref(cc),
                                               _ptr(trace), trace_cap)
        return int(attempts), int(bc.value), int(cc.value), trace[:min(attempts, trace_cap)]


def ref_available() -> bool:
    return os.path.exists(REF_SO) or os.path.isdir(REFERENCE_ROOT)
