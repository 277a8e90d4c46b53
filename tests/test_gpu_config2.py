"""SURVEY.md section 8(d), config 2 exactly as written, and the paths round 1 left unasserted:

  (i)   top-k pop sequences bit-exact for ALL positions of the first 64 KiB + 4096 uniformly sampled positions of
        the 1 MiB mixed corpus, under the init state and under states snapshotted along an ANNEALED slab;
  (ii)  mg_score_slabs against the oracle on >= 64 slabs (all-literal, greedy, annealed snapshots, random valid
        mutations) - exact u64;
  (iii) the exact early exit fires (stats["rejoined"] > 0) and leaves the trace unchanged;
  (iv)  a clock-boxed run on a zero-run binary makes the match finder give a bucket up (finder_gave_up > 0) and
        the trajectories still equal the oracle's;
  (v)   >= 1024 chains x >= 50 evaluations at 64 KiB against the oracle on a sampled subset of chains;
  (vi)  the device-built bigram index against an independent restatement of memoize_bigram_positions.

Reference: src/top_k_packet_finder.c:95-138, src/packet_slab_neighbour.c:154-173, src/substring_enumerator.c:26-47.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MIB = 1 << 20


def same_packets(a, b):
    return (a["type"] == b["type"]).all() and (a["dist"] == b["dist"]).all() and (a["len"] == b["len"]).all()


def boundaries(slab):
    out, p = [], 0
    lens = slab["len"].astype(np.int64)
    n = slab.size
    while p < n:
        out.append(p)
        p += int(lens[p])
    return np.array(out, dtype=np.uint64)


@pytest.fixture(scope="module")
def mg():
    import megalania_b200 as m
    m.load_library()
    return m


@pytest.fixture(scope="module")
def big(corpora):
    return corpora("mixed", MIB)


@pytest.fixture(scope="module")
def big_ctx(mg, big):
    ctx = mg.Context(big)
    yield ctx
    ctx.close()


@pytest.fixture(scope="module")
def big_greedy(port, big):
    return port.greedy_slab(big)


@pytest.fixture(scope="module")
def annealed(mg, big_ctx, big_greedy):
    """Snapshots along annealing runs of the 1 MiB corpus: chains started from the greedy parse and from the
    all-literal slab, current and best slabs."""
    snaps = []
    an = mg.Annealer(big_ctx, 24, seed=20261018)
    an.set_slab(big_greedy)
    an.run(40, step=1)
    for c in range(24):
        snaps.append(an.get_slab(c))
    for c in range(0, 24, 3):
        snaps.append(an.get_slab(c, best=True))
    an.set_slab(None)
    an.run(12)
    for c in range(12):
        snaps.append(an.get_slab(c))
    an.close()
    return snaps


def test_config2_topk_first_64k_plus_4096_sampled_init_state(mg, port, big, big_ctx):
    lit = mg.literal_slab(MIB)
    rng = np.random.default_rng(2026)
    pos = np.concatenate([np.arange(65536), 65536 + np.sort(rng.choice(MIB - 65536, 4096, replace=False))]).astype(np.uint64)
    pops, prices, counts = big_ctx.find_topk(lit, pos, state_mode=0)
    wp, wprice, wc = port.topk_many_priced(big, lit, 0, pos)
    assert (counts == wc).all()
    assert same_packets(pops, wp)
    assert (prices == wprice).all()


def test_config2_topk_states_along_an_annealed_slab(mg, port, big, big_ctx, annealed):
    rng = np.random.default_rng(7)
    for slab in (annealed[0], annealed[-1]):  # one matured from the greedy parse, one young from the all-literal slab
        assert port.slab_valid(big, slab)
        b = boundaries(slab)
        head = b[b < 65536]
        tail = b[b >= 65536]
        sample = np.concatenate([head, np.sort(rng.choice(tail, min(4096, tail.size), replace=False))]).astype(np.uint64)
        pops, prices, counts = big_ctx.find_topk(slab, sample, state_mode=1)
        wp, wprice, wc = port.topk_many_priced(big, slab, 1, sample)
        assert (counts == wc).all()
        assert same_packets(pops, wp)
        assert (prices == wprice).all()


def shrink_mutations(slab, rng, count):
    """Random valid mutations that keep every later packet decodable: a MATCH loses its last byte, which
    becomes a LITERAL (the rep-distance history is unchanged)."""
    out = slab.copy()
    b = boundaries(out)
    cand = b[(out["type"][b.astype(np.int64)] == 2) & (out["len"][b.astype(np.int64)] > 2)]
    for x in rng.choice(cand, min(count, cand.size), replace=False):
        x = int(x)
        ln = int(out[x]["len"])
        out[x] = (2, int(out[x]["dist"]), ln - 1)
        out[x + ln - 1] = (1, 0, 1)
    return out


def test_config2_score_64_slabs(mg, port, big, big_ctx, big_greedy, annealed):
    rng = np.random.default_rng(99)
    slabs = [mg.literal_slab(MIB), big_greedy] + list(annealed)
    while len(slabs) < 66:
        base = annealed[int(rng.integers(len(annealed)))] if len(slabs) % 2 else big_greedy
        slabs.append(shrink_mutations(base, rng, int(rng.integers(1, 4000))))
    assert len(slabs) >= 64
    want = [port.slab_cost(big, s) for s in slabs]
    # in batches, so that the host copy stays small
    got = []
    for i in range(0, len(slabs), 16):
        got += [int(x) for x in big_ctx.score_slabs(np.concatenate(slabs[i:i + 16]))]
    assert got == want
    assert len(set(want)) >= 40  # the slabs really differ (chains that accepted nothing keep their start slab)


def test_exact_early_exit_fires_and_changes_nothing(mg, port, corpora):
    n, chains, evals, seed = 65536, 4, 120, 808
    data = corpora("text", n)
    init = port.greedy_slab(data)
    traces = {}
    with mg.Context(data) as ctx:
        for early in (True, False):
            an = mg.Annealer(ctx, chains, trace_capacity=evals * 64 + 1024, seed=seed, checkpoint_stride=512)
            an.set_slab(init, adopt_cost=False)
            st = an.run(evals, step=1, early_exit=early)
            assert st["evals"] == chains * evals
            if early:
                assert st["rejoined"] > 0, "no proposal re-joined the current slab's checkpoints: the early exit is untested"
                packets_early = st["packets_scored"]
            else:
                assert st["rejoined"] == 0
                assert st["packets_scored"] > packets_early
            traces[early] = [an.trace(c).copy() for c in range(chains)]
            cur, best = an.costs()
            slabs = [an.get_slab(c) for c in range(chains)]
            an.close()
            for c in range(chains):
                assert ctx.score_slab(slabs[c]) == int(cur[c])
        for c in range(chains):
            a, b = traces[True][c], traces[False][c]
            assert len(a) == len(b) and (a["cost"] == b["cost"]).all() and (a["flags"] == b["flags"]).all()
            assert (a["undo_count"] == b["undo_count"]).all()
        for c in (0, chains - 1):
            slab, bslab = init.copy(), init.copy()
            attempts, bc, cc, _, trace = port.anneal_epoch(data, slab, bslab, 0, 0, rng_mode=1,
                                                           rng_state=port.chain_seed(seed, c), step=1, evals=evals)
            g = traces[True][c]
            assert len(g) == attempts and (g["cost"] == trace["cost"]).all() and (g["flags"] == trace["flags"]).all()


def test_clock_boxed_finder_gives_up_and_redraws(mg, port):
    """Zero runs make one bucket of ~16 k occurrences; a tight clock box forces FIND_GAVE_UP (the proposal is
    taken back whole and drawn again by the next launch): trajectories must not notice."""
    n, chains, seed = 32768, 8, 4711
    rng = np.random.default_rng(3)
    blob = bytearray(n)
    # islands of random bytes in a sea of zeros
    for start in range(0, n, 2048):
        blob[start:start + 64] = rng.integers(1, 256, 64, dtype=np.uint8).tobytes()
    data = bytes(blob)
    lit = mg.literal_slab(n)
    with mg.Context(data) as ctx:
        an = mg.Annealer(ctx, chains, trace_capacity=4096, seed=seed, checkpoint_stride=512)
        an.set_slab(None, adopt_cost=False)
        got = [[] for _ in range(chains)]
        gave_up = 0
        for _ in range(60):
            st = an.run(100000, cycle_budget=400_000, suspend=True, first_eval=mg.CONTINUE_EVALS)
            gave_up += st["finder_gave_up"]
            for c in range(chains):
                got[c].append(an.trace(c))
        assert gave_up > 0, "the deadline never passed inside the match finder: FIND_GAVE_UP is untested"
        cur, best = an.costs()
        for c in range(chains):
            g = np.concatenate(got[c])
            evals = int((g["flags"] & 1).sum())
            assert evals >= 3
            slab, bslab = lit.copy(), lit.copy()
            attempts, bc, cc, _, trace = port.anneal_epoch(data, slab, bslab, 0, 0, rng_mode=1,
                                                           rng_state=port.chain_seed(seed, c), evals=evals)
            assert len(g) >= attempts
            assert (g["flags"][:attempts] == trace["flags"]).all()
            assert (g["cost"][:attempts] == trace["cost"]).all()
            assert int(best[c]) == bc
        an.close()


def test_1024_chains_50_evals_64k_sampled_against_oracle(mg, port, corpora):
    n, chains, evals, seed = 65536, 1024, 50, 1234567
    data = corpora("mixed", n)
    lit = mg.literal_slab(n)
    with mg.Context(data) as ctx:
        an = mg.Annealer(ctx, chains, trace_capacity=1024, seed=seed)
        an.set_slab(None)
        # in time-boxed, suspended steps like the bench: the population shape the headline is measured on
        traces = {c: [] for c in (0, 1, 31, 32, 333, 512, 777, 1023)}
        total = 0
        for _ in range(90):
            st = an.run(1000, cycle_budget=6_000_000, suspend=True, first_eval=mg.CONTINUE_EVALS)
            assert st["log_overflows"] == 0
            total += st["evals"]
            for c in traces:
                traces[c].append(an.trace(c))
        assert total >= chains * evals
        cur, best = an.costs()
        for c, parts in traces.items():
            g = np.concatenate(parts)
            ev = int((g["flags"] & 1).sum())
            assert ev >= evals
            slab, bslab = lit.copy(), lit.copy()
            attempts, bc, cc, _, trace = port.anneal_epoch(data, slab, bslab, 0, 0, rng_mode=1,
                                                           rng_state=port.chain_seed(seed, c), evals=ev)
            assert len(g) >= attempts
            assert (g["flags"][:attempts] == trace["flags"]).all()
            assert (g["cost"][:attempts] == trace["cost"]).all()
            assert (g["undo_count"][:attempts] == trace["undo_count"]).all()
            assert int(best[c]) == bc and int(cur[c]) == cc
            assert same_packets(an.get_slab(c), slab)
        an.close()


@pytest.mark.parametrize("kind,n", [("mixed", 65536), ("binary", 4096), ("text", 1000)])
def test_device_bigram_index(mg, corpora, port, kind, n):
    """K1 against memoize_bigram_positions: per bucket the ascending list of positions whose two bytes are the key."""
    data = corpora(kind, n)
    d = np.frombuffer(data, dtype=np.uint8).astype(np.uint32)
    keys = (d[:-1] << 8) | d[1:]
    order = np.argsort(keys, kind="stable").astype(np.uint32)
    counts = np.bincount(keys, minlength=65536)
    want_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    with mg.Context(data) as ctx:
        start, occ = ctx.bigram_index()
    assert (start == want_start).all()
    assert (occ == order).all()
    # ... and the walk over it reproduces the reference's callback counts at a few positions
    for pos in (1, n // 3, n - 2):
        k = int(keys[pos])
        earlier = occ[start[k]:start[k + 1]]
        earlier = earlier[earlier < pos]
        total = 0
        for o in earlier:
            ln = 0
            while pos + ln < n and ln < 273 and data[int(o) + ln] == data[pos + ln]:
                ln += 1
            total += max(0, ln - 1)
        assert total == port.substring_count(data, pos)


@pytest.mark.parametrize("window,max_occ", [(1500, 0), (0, 40), (3000, 64)])
def test_finder_limits_match_the_restricted_oracle(mg, port, corpora, window, max_occ):
    """a real window limit (src/substring_enumerator.c:97 is commented out in the reference) and a cap
    on the occurrences walked.  With limits the lists are the reference's lists restricted to the surviving
    occurrences; whole annealing trajectories under limits equal the oracle's under the same limits."""
    n = 16384
    data = corpora("mixed", n)
    lit = mg.literal_slab(n)
    try:
        port.set_finder_limits(window, max_occ)
        with mg.Context(data) as ctx:
            ctx.set_finder_limits(window, max_occ)
            pos = np.arange(n, dtype=np.uint64)
            pops, prices, counts = ctx.find_topk(lit, pos, state_mode=0)
            wp, wprice, wc = port.topk_many_priced(data, lit, 0, pos)
            assert (counts == wc).all() and same_packets(pops, wp) and (prices == wprice).all()
            chains, evals, seed = 3, 120, 55
            an = mg.Annealer(ctx, chains, trace_capacity=evals * 64 + 1024, seed=seed, checkpoint_stride=512)
            an.set_slab(None)
            an.run(evals)
            cur, best = an.costs()
            for c in range(chains):
                slab, bslab = lit.copy(), lit.copy()
                attempts, bc, cc, _, trace = port.anneal_epoch(data, slab, bslab, 0, 0, rng_mode=1,
                                                               rng_state=port.chain_seed(seed, c), evals=evals)
                got = an.trace(c)
                assert len(got) == attempts and (got["cost"] == trace["cost"]).all() and (got["flags"] == trace["flags"]).all()
                assert int(cur[c]) == cc and int(best[c]) == bc
            an.close()
            # the limits really bite: the unrestricted lists differ somewhere
            ctx.set_finder_limits(0, 0)
            pops0, prices0, counts0 = ctx.find_topk(lit, pos, state_mode=0)
            assert not ((counts0 == counts).all() and (prices0 == prices).all())
    finally:
        port.set_finder_limits(0, 0)


def test_literal_queue_edge_windows(mg, port):
    """The per-lane literal queues (walk_windows): runs of every byte value (some put two hot slots of the
    tree into one bank: more than 16, some more than 24 rounds -> third block / single-step fallback),
    alternating bytes, random bytes, windows interrupted by matches - cost and the whole model after
    several prefixes must equal the oracle's."""
    rng = np.random.default_rng(99)
    parts = []
    for v in range(256):
        parts.append(bytes([v]) * 40)
    for a, b in ((0x00, 0xff), (0x65, 0x30), (0x41, 0xc1), (0x18, 0x98)):
        parts.append(bytes([a, b]) * 48)
    parts.append(rng.integers(0, 256, 4096, dtype=np.uint8).tobytes())
    parts.append(rng.integers(0x60, 0x68, 2048, dtype=np.uint8).tobytes())
    data = b"".join(parts)
    data = data[: len(data) // 32 * 32 + 7]  # a ragged tail window
    n = len(data)
    lit = mg.literal_slab(n)
    greedy = port.greedy_slab(data)
    mixed = lit.copy()
    for p in range(300, n - 64, 517):  # sparse matches: most windows stay whole, some are cut
        if greedy[p]["type"] == mg.MATCH:  # a MATCH is valid wherever its bytes match, whatever came before
            mixed[p] = greedy[p]
    with mg.Context(data) as ctx:
        costs = ctx.score_slabs(np.concatenate([lit, greedy]))
        assert int(costs[0]) == port.slab_cost(data, lit)
        assert int(costs[1]) == port.slab_cost(data, greedy)
        for stop in (0, 32, 40 * 7, 40 * 0x65 + 32, 40 * 256, n // 2 // 32 * 32, n):
            got = ctx.model_after_prefix(lit, stop)
            want = port.model_after_prefix(data, lit, stop)
            assert got.tobytes() == want.tobytes(), stop
        assert ctx.score_slab(mixed) == port.slab_cost(data, mixed)


@pytest.mark.parametrize("kind,n", [("text", 4096), ("binary", 4096)])
def test_greedy_init_equals_oracle(mg, port, corpora, kind, n):
    """mg_anneal_greedy_init with one region and no finder limits is the oracle's greedy parse packet for packet
    (exact top-k at every packet, cheapest candidate per byte, adaptive model); with many regions the stitched,
    repaired slab is valid, priced exactly and no worse than the all-literal slab."""
    data = corpora(kind, n)
    want = port.greedy_slab(data)
    with mg.Context(data) as ctx:
        an = mg.Annealer(ctx, 4, seed=3)
        an.set_slab(None)
        cost = an.greedy_init(1, 0)
        got = an.get_slab(0)
        b = 0
        while b < n:  # compare the live chain: slots behind a packet are dead
            assert (int(got[b]["type"]), int(got[b]["dist"]), int(got[b]["len"])) == \
                   (int(want[b]["type"]), int(want[b]["dist"]), int(want[b]["len"])), b
            b += int(want[b]["len"])
        assert cost == port.slab_cost(data, want)
        cost16 = an.greedy_init(16, 1)
        slab16 = an.get_slab(1)
        assert cost16 == port.slab_cost(data, slab16) == ctx.score_slab(slab16)
        assert cost16 < port.slab_cost(data, mg.literal_slab(n))
        stream = ctx.encode_slab(slab16)
        import lzma
        assert lzma.decompress(stream, format=lzma.FORMAT_ALONE) == data
        an.broadcast_chain(1)
        cur, _ = an.costs()
        assert (cur == cost16).all()
        an.close()
