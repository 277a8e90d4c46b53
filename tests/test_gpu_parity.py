"""GPU parity tests proper: the CUDA path, through the C ABI, against the oracle on the same
seeded inputs.  Bit-exact everywhere (all arithmetic on this path is integer)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HELLO = b"hello hello"


def boundaries(slab):
    out, p = [], 0
    while p < slab.size:
        out.append(p)
        p += int(slab[p]["len"])
    return np.array(out, dtype=np.uint64)


def same_packets(a, b):
    return (a["type"] == b["type"]).all() and (a["dist"] == b["dist"]).all() and (a["len"] == b["len"]).all()


@pytest.fixture(scope="module")
def mg():
    import megalania_b200 as m
    m.load_library()
    return m


def test_hello_known_answers(mg, port):
    # SURVEY.md Appendix B, recorded from the reference binary
    with mg.Context(HELLO) as ctx:
        lit = mg.literal_slab(11)
        assert ctx.score_slab(lit) == 182379
        s2 = lit.copy()
        s2[6] = (mg.MATCH, 5, 5)
        assert ctx.score_slab(s2) == 132981
        pops, prices, counts = ctx.find_topk(lit, np.arange(11), state_mode=0)
        expect = {3: [(3, 0, 1)], 6: [(2, 5, 2), (2, 5, 3), (2, 5, 4), (2, 5, 5)], 7: [(2, 5, 2), (2, 5, 3), (2, 5, 4)],
                  8: [(2, 5, 2), (2, 5, 3)], 9: [(2, 5, 2), (3, 0, 1)]}
        for i in range(11):
            got = [(int(p["type"]), int(p["dist"]), int(p["len"])) for p in pops[i][:counts[i]]]
            assert got == expect.get(i, []), (i, got)
        assert ctx.encode_slab(s2) == port.encode_slab(HELLO, s2)


@pytest.mark.parametrize("kind,n", [("text", 4096), ("binary", 4096), ("mixed", 16384)])
def test_cost_model_bytes(mg, port, corpora, kind, n):
    data = corpora(kind, n)
    lit = mg.literal_slab(n)
    greedy = port.greedy_slab(data)
    with mg.Context(data) as ctx:
        slabs = np.concatenate([lit, greedy])
        costs = ctx.score_slabs(slabs)
        assert int(costs[0]) == port.slab_cost(data, lit)
        assert int(costs[1]) == port.slab_cost(data, greedy)
        b = boundaries(greedy)
        for stop in (0, int(b[len(b) // 3]), int(b[-1]), n):
            got = ctx.model_after_prefix(greedy, stop)
            want = port.model_after_prefix(data, greedy, stop)
            assert got.tobytes() == want.tobytes(), stop
        assert ctx.encode_slab(greedy) == port.encode_slab(data, greedy)
        assert ctx.encode_slab_buffer(lit) == port.encode_slab(data, lit)


@pytest.mark.parametrize("kind,n", [("text", 4096), ("binary", 4096), ("mixed", 16384)])
def test_topk_all_positions(mg, port, corpora, kind, n):
    data = corpora(kind, n)
    lit = mg.literal_slab(n)
    greedy = port.greedy_slab(data)
    with mg.Context(data) as ctx:
        pos = np.arange(n, dtype=np.uint64)
        pops, prices, counts = ctx.find_topk(lit, pos, state_mode=0)
        wp, wprice, wc = port.topk_many_priced(data, lit, 0, pos)
        assert (counts == wc).all()
        assert same_packets(pops, wp)
        assert (prices == wprice).all()
        b = boundaries(greedy)
        pops, prices, counts = ctx.find_topk(greedy, b, state_mode=1)
        wp, wprice, wc = port.topk_many_priced(data, greedy, 1, b)
        assert (counts == wc).all()
        assert same_packets(pops, wp)
        assert (prices == wprice).all()


@pytest.mark.parametrize("kind,n,step", [("text", 2048, 0), ("binary", 2048, 0), ("mixed", 4096, 1), ("text", 8192, 1)])
def test_anneal_trace_matches_oracle(mg, port, corpora, kind, n, step):
    """Same generator, same schedule: every proposal's cost, decision and edit count must agree."""
    data = corpora(kind, n)
    chains, evals, seed = 4, 200, 99
    init = mg.literal_slab(n) if step == 0 else port.greedy_slab(data)
    with mg.Context(data) as ctx:
        an = mg.Annealer(ctx, chains, trace_capacity=evals * 64 + 1024, seed=seed, checkpoint_stride=512)
        an.set_slab(init if step else None, adopt_cost=False)
        stats = an.run(evals, step=step)
        cur, best = an.costs()
        assert stats["evals"] == chains * evals
        assert stats["log_overflows"] == 0
        for c in range(chains):
            slab, bslab = init.copy(), init.copy()
            attempts, bc, cc, _, trace = port.anneal_epoch(data, slab, bslab, 0, 0, rng_mode=1,
                                                           rng_state=port.chain_seed(seed, c), step=step, evals=evals)
            got = an.trace(c)
            assert len(got) == attempts
            assert (got["flags"] == trace["flags"]).all()
            assert (got["cost"] == trace["cost"]).all()
            assert (got["undo_count"] == trace["undo_count"]).all()
            assert int(cur[c]) == cc and int(best[c]) == bc
            assert same_packets(an.get_slab(c), slab)
            assert same_packets(an.get_slab(c, best=True), bslab)
            assert ctx.score_slab(an.get_slab(c)) == cc
        an.close()


@pytest.mark.parametrize("kind,n,step,budget", [("text", 8192, 0, 3000), ("mixed", 8192, 1, 1500), ("binary", 4096, 1, 700)])
def test_suspended_proposals_do_not_change_trajectories(mg, port, corpora, kind, n, step, budget):
    """Exact packet budgets: a proposal cut at a checkpoint and carried on by the next launch must
    leave every chain on the trajectory the oracle follows without any cuts."""
    data = corpora(kind, n)
    chains, seed, launches = 4, 31, 40
    init = mg.literal_slab(n) if step == 0 else port.greedy_slab(data)
    with mg.Context(data) as ctx:
        an = mg.Annealer(ctx, chains, trace_capacity=4096, seed=seed, checkpoint_stride=512)
        an.set_slab(init if step else None, adopt_cost=False)
        got = [[] for _ in range(chains)]
        packets = 0
        for _ in range(launches):
            st = an.run(1000, step=step, packet_budget=budget, suspend=True, first_eval=mg.CONTINUE_EVALS)
            packets += st["packets_scored"]
            for c in range(chains):
                got[c].append(an.trace(c))
        # the budget is exact to within the distance to the next checkpoint
        assert packets <= chains * launches * (budget + 2 * 512 + 8)
        assert packets >= chains * launches * budget
        cur, best = an.costs()
        for c in range(chains):
            g = np.concatenate(got[c])
            evals = int((g["flags"] & 1).sum())
            assert evals >= 10
            slab, bslab = init.copy(), init.copy()
            attempts, bc, cc, _, trace = port.anneal_epoch(data, slab, bslab, 0, 0, rng_mode=1,
                                                           rng_state=port.chain_seed(seed, c), step=step, evals=evals)
            # the GPU may have drawn failed proposals after its last success
            assert len(g) >= attempts
            assert (g["flags"][:attempts] == trace["flags"]).all()
            assert (g["cost"][:attempts] == trace["cost"]).all()
            assert (g["undo_count"][:attempts] == trace["undo_count"]).all()
            assert (g["flags"][attempts:] == 0).all()
            assert int(best[c]) == bc
        an.close()


def test_clock_boxed_steps_keep_trajectories(mg, port, corpora):
    """SM-clock budgets (mg_anneal_run_params.cycle_budget): how many evaluations fit in a step is
    not reproducible, the trajectory of every chain is - suspended proposals included."""
    n, chains, seed = 16384, 6, 77
    data = corpora("mixed", n)
    with mg.Context(data) as ctx:
        an = mg.Annealer(ctx, chains, trace_capacity=8192, seed=seed, checkpoint_stride=512)
        an.set_slab(None, adopt_cost=False)
        got = [[] for _ in range(chains)]
        for _ in range(12):
            st = an.run(100000, cycle_budget=3_000_000, suspend=True, first_eval=mg.CONTINUE_EVALS)
            assert st["max_chain_cycles"] < 100_000_000  # a chain stops at its next checkpoint or evaluation boundary
            for c in range(chains):
                got[c].append(an.trace(c))
        cur, best = an.costs()
        lit = mg.literal_slab(n)
        for c in range(chains):
            g = np.concatenate(got[c])
            evals = int((g["flags"] & 1).sum())
            assert evals >= 5
            slab, bslab = lit.copy(), lit.copy()
            attempts, bc, cc, _, trace = port.anneal_epoch(data, slab, bslab, 0, 0, rng_mode=1,
                                                           rng_state=port.chain_seed(seed, c), evals=evals)
            assert len(g) >= attempts
            assert (g["flags"][:attempts] == trace["flags"]).all()
            assert (g["cost"][:attempts] == trace["cost"]).all()
            assert int(best[c]) == bc
        an.close()
