"""The oracle port against the unmodified reference compiled into oracle/_ref, on fresh seeded
inputs (skipped where neither oracle/_ref nor /root/reference exists)."""
import numpy as np
import pytest

from oracle import oracle_lib as ol
from tools import corpus


def boundaries(slab):
    out, p = [], 0
    while p < slab.size:
        out.append(p)
        p += int(slab[p]["len"])
    return out


def test_struct_sizes(ref):
    # SURVEY.md Appendix B
    assert ref.lib.mgref_sizeof_packet() == 12 == ol.PACKET_DTYPE.itemsize
    assert ref.lib.mgref_sizeof_state() == 5280


def test_tables_and_rand(port, ref):
    assert (port.price_table().astype(np.uint64) == ref.price_table()).all()
    for seed in (1, 42, 1673551):
        assert (port.rand_stream(seed, 2000) == ref.rand_stream(seed, 2000)).all()


def test_heap_ties(port, ref):
    rng = corpus.SplitMix64(7)
    for trial in range(50):
        count = 1 + rng.below(300)
        k = 1 + rng.below(32)
        keys = [int(rng.below(1 + trial % 9)) for _ in range(count)]
        assert (port.heap_topk(keys, k) == ref.heap_topk(keys, k)).all()


@pytest.mark.parametrize("kind,n,seed", [("text", 3000, 1), ("binary", 3000, 2), ("mixed", 6000, 3)])
def test_three_parity_functions(port, ref, kind, n, seed):
    data = corpus.make(kind, n, seed)
    lit = ol.literal_slab(n)
    greedy = port.greedy_slab(data)
    for slab in (lit, greedy):
        assert port.slab_cost(data, slab) == ref.slab_cost(data, slab)
        assert port.encode_slab(data, slab) == ref.encode_slab(data, slab)
    b = boundaries(greedy)
    for stop in (b[1], b[len(b) // 2], b[-1]):
        assert port.prefix_cost(data, greedy, stop) == ref.prefix_cost(data, greedy, stop)
        assert port.model_after_prefix(data, greedy, stop).tobytes() == ref.model_after_prefix(data, greedy, stop).tobytes()
    pos = np.arange(n)
    pp, pc = port.topk_many(data, lit, 0, pos)
    rp, rc = ref.topk_many(data, lit, 0, pos)
    assert (pc == rc).all() and (pp == rp).all()
    pp, pc = port.topk_many(data, greedy, 1, b)
    rp, rc = ref.topk_many(data, greedy, 1, b)
    assert (pc == rc).all() and (pp == rp).all()


@pytest.mark.parametrize("kind,n,step", [("text", 1500, 0), ("binary", 1500, 0), ("mixed", 3000, 1), ("text", 3000, 2)])
def test_anneal_epoch_trajectory(port, ref, kind, n, step):
    data = corpus.make(kind, n, 11)
    start = ol.literal_slab(n) if step == 0 else port.greedy_slab(data)
    s1, s2, b1, b2 = start.copy(), start.copy(), start.copy(), start.copy()
    a1 = port.anneal_epoch(data, s1, b1, 0, 0, rng_mode=0, step=step, evals=400)
    a2 = ref.anneal_epoch(data, s2, b2, 0, 0, step=step, evals=400)
    assert a1[:3] == a2[:3]
    assert (a1[4] == a2[3]).all()
    assert (s1 == s2).all() and (b1 == b2).all()


def test_edge_inputs(port, ref):
    for data in (b"a", b"ab", b"aaaa", b"abcabcabcabc", bytes(300), bytes(range(256))):
        n = len(data)
        lit = ol.literal_slab(n)
        assert port.slab_cost(data, lit) == ref.slab_cost(data, lit)
        assert port.encode_slab(data, lit) == ref.encode_slab(data, lit)
        pp, pc = port.topk_many(data, lit, 0, np.arange(n))
        rp, rc = ref.topk_many(data, lit, 0, np.arange(n))
        assert (pc == rc).all() and (pp == rp).all()
