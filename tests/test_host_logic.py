"""Host-side pieces that need no GPU: corpus generators, the reciprocal-multiply division the
finder uses, argument checking of the Python mirror."""
import hashlib

import numpy as np
import pytest

from tools import corpus


def test_corpus_is_deterministic_and_shaped():
    a = corpus.make("mixed", 1 << 16)
    b = corpus.make("mixed", 1 << 16)
    assert a == b and len(a) == 1 << 16
    assert corpus.text(1000, 1) != corpus.text(1000, 2)
    text = corpus.text(1 << 14)
    assert sum(c in b" .,?\n" or chr(c).isalpha() for c in text) == len(text)
    binary = corpus.binary(1 << 14)
    zero_frac = binary.count(0) / len(binary)
    assert 0.2 < zero_frac < 0.8
    big = corpus.corpus16(1 << 18)
    assert len(big) == 1 << 18
    # pinned digest: the generators feed every golden vector
    assert hashlib.sha256(corpus.text(4096)).hexdigest()[:16] == hashlib.sha256(corpus.make("text", 4096, 12345)).hexdigest()[:16]


def test_reciprocal_division_is_exact():
    """mg_finder.cuh per_byte(): floor(c/len) == umulhi(c, floor((2^32-1)/len)+1) whenever c*len < 2^32."""
    rng = np.random.default_rng(0)
    for length in range(2, 274):
        m = 0xFFFFFFFF // length + 1
        limit = min((1 << 32) // length, 1 << 22)
        cs = np.concatenate([rng.integers(0, limit, 2000), np.arange(0, 600), np.array([limit - 1, limit - 2])]).astype(np.uint64)
        got = (cs * np.uint64(m)) >> np.uint64(32)
        assert (got == cs // np.uint64(length)).all(), length


def test_python_mirror_checks_arguments():
    import megalania_b200 as mg
    from megalania_b200 import api
    slab = mg.literal_slab(8)
    assert slab.dtype.itemsize == 12
    with pytest.raises(ValueError):
        api._slab_ptr(slab, 9)
    repacked = np.concatenate([slab, slab])
    fixed = api.as_slab(repacked)
    assert fixed.dtype == api.PACKET_DTYPE and fixed.size == 16 and (fixed["len"] == 1).all()


def test_region_plan_covers_the_input_and_balances_groups():
    from megalania_b200.cooperative import region_plan
    for n, chains, group, shift in ((65536, 4736, 8, 12345), (4096, 4736, 8, 7), (1 << 20, 4736, 1, 999999), (300, 64, 4, 5)):
        bounds, region_of_chain = region_plan(n, chains, group, shift)
        assert bounds[0] == 0 and bounds[-1] == n
        assert (np.diff(bounds.astype(np.int64)) > 0).all()
        nreg = bounds.size - 1
        assert region_of_chain.min() == 0 and region_of_chain.max() == nreg - 1
        counts = np.bincount(region_of_chain, minlength=nreg)
        assert counts.min() >= 1 and counts.max() - counts.min() <= 1
    # boundaries move with the shift
    a, _ = region_plan(65536, 4736, 8, 0)
    b, _ = region_plan(65536, 4736, 8, 50)
    assert not np.array_equal(a, b)


def test_distributed_plan_gives_every_rank_its_own_regions():
    from megalania_b200.cooperative import distributed_plan
    for n, chains, world, group in ((1 << 20, 4736, 8, 8), (65536, 4736, 2, 8), (4096, 64, 4, 4)):
        bounds, owner_rank, regions_of_chains = distributed_plan(n, chains, world, group, 4242)
        nreg = bounds.size - 1
        assert bounds[0] == 0 and bounds[-1] == n and (np.diff(bounds.astype(np.int64)) > 0).all()
        assert owner_rank.size == nreg and set(owner_rank.tolist()) <= set(range(world))
        seen = np.zeros(nreg, dtype=bool)
        for rank in range(world):
            mine = regions_of_chains(rank)
            assert mine.size == chains
            assert (owner_rank[mine] == rank).all()      # a rank only anneals regions it owns
            seen[np.unique(mine)] = True
        assert seen.all()                                 # every region is annealed by someone


def test_c_temper_rule_matches_the_python_rule():
    """mg_temper_decide (the swap rule the multi-GPU C host uses, no device needed) against
    megalania_b200.tempering.exchange_temperatures on random ladders, both parities, ties included."""
    import numpy as np
    import megalania_b200 as mg
    from megalania_b200 import api, tempering
    mg.load_library()
    rng = np.random.default_rng(5)
    for trial in range(40):
        count = int(rng.integers(1, 400))
        temps = tempering.temperature_ladder(count, 10.0, 1e6)[rng.permutation(count)]
        if trial % 5 == 0 and count > 3:
            temps[1] = temps[0]          # equal temperatures: stable order decides
            temps[2] = 0.0               # a frozen replica never swaps
        costs = rng.integers(1_000_000, 9_000_000_000, count).astype(np.uint64)
        for rnd in (trial, trial + 1):
            want = tempering.exchange_temperatures(costs.astype(np.int64), temps, rnd, seed=trial * 7919)
            got = api.temper_decide(costs, temps, rnd, seed=trial * 7919)
            assert (got == want).all(), (trial, rnd)
            assert sorted(got.tolist()) == sorted(temps.tolist())
