"""Cooperative regions on the GPU: every cost the merge reports must be the oracle's cost of the very
slab it produced, the slab must decode to the input, and rounds never lose ground at temperature 0."""
import lzma

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mg():
    import megalania_b200 as m
    m.load_library()
    return m


def live_starts(slab):
    out, p = [], 0
    while p < slab.size:
        out.append(p)
        p += int(slab[p]["len"])
    return np.array(out)


@pytest.mark.parametrize("kind,n,start", [("text", 8192, "literal"), ("binary", 4096, "greedy"), ("mixed", 16384, "greedy")])
def test_merge_is_exact_and_monotone(mg, port, corpora, kind, n, start):
    from megalania_b200.cooperative import CooperativeAnnealer
    data = corpora(kind, n)
    init = None if start == "literal" else port.greedy_slab(data)
    with mg.Context(data) as ctx:
        an = mg.Annealer(ctx, 256, seed=9, checkpoint_stride=512)
        coop = CooperativeAnnealer(an, group=4, min_region=64, seed=1)
        cost0 = coop.start(init)
        assert cost0 == port.slab_cost(data, init if init is not None else mg.literal_slab(n))
        last = cost0
        for _ in range(6):
            rec = coop.round(packet_budget=40 * n, evals=400)
            slab = coop.slab()
            # the cost the device reports is the oracle's cost of the slab the device built
            assert rec["cost"] == port.slab_cost(data, slab)
            assert rec["cost"] <= last
            last = rec["cost"]
            # every chain now holds that slab, priced
            cur, _ = an.costs()
            assert (cur == rec["cost"]).all()
        assert last < cost0
        assert any(r["kept"] == "merged" for r in coop.history)
        stream = ctx.encode_slab(coop.slab())
        assert stream == port.encode_slab(data, coop.slab())
        assert lzma.decompress(stream, format=lzma.FORMAT_ALONE) == data
        an.close()


def test_region_confined_chains_only_start_edits_inside(mg, port, corpora):
    """A confined chain's first edit of every accepted proposal lies in its range (later edits are the
    repairs the reference makes downstream of a mutation)."""
    n = 8192
    data = corpora("text", n)
    greedy = port.greedy_slab(data)
    with mg.Context(data) as ctx:
        an = mg.Annealer(ctx, 8, seed=4, checkpoint_stride=512)
        an.set_slab(greedy)
        regions = np.array([[c * 1024, (c + 1) * 1024] for c in range(8)], dtype=np.uint32)
        temps = np.zeros(8, dtype=np.float32)
        st = an.run(300, schedule=mg.SCHEDULE_TEMPERATURE, temperatures=temps, regions=regions)
        assert st["evals"] == 8 * 300
        moved = 0
        for c in range(8):
            slab = an.get_slab(c)
            changed = np.nonzero((slab["type"] != greedy["type"]) | (slab["dist"] != greedy["dist"]) |
                                 (slab["len"] != greedy["len"]))[0]
            if changed.size:
                moved += 1
                assert changed.min() >= regions[c, 0]
            assert ctx.score_slab(slab) == port.slab_cost(data, slab)
        assert moved >= 4  # at temperature 0 a chain only keeps moves that do not raise the cost
        an.close()


def test_merge_in_two_halves_equals_the_one_call(mg, port, corpora):
    """mg_anneal_merge_export + mg_anneal_merge_import (the multi-GPU halves, with the parts of two 'processes'
    summed in between) build the same slab at the same cost as mg_anneal_merge_regions."""
    import torch
    from megalania_b200.cooperative import region_plan
    n = 8192
    data = corpora("mixed", n)
    greedy = port.greedy_slab(data)
    with mg.Context(data) as ctx:
        an = mg.Annealer(ctx, 64, seed=12, checkpoint_stride=512)
        an.set_slab(greedy)
        bounds, region_of_chain = region_plan(n, 64, 4, 1234)
        nreg = bounds.size - 1
        regions = np.stack([bounds[region_of_chain], bounds[region_of_chain + 1]], axis=1).astype(np.uint32)
        an.run(200, schedule=mg.SCHEDULE_TEMPERATURE, temperatures=np.zeros(64, dtype=np.float32), regions=regions)
        cur, _ = an.costs()
        owners = np.array([np.nonzero(region_of_chain == r)[0][np.argmin(cur[region_of_chain == r])] for r in range(nreg)],
                          dtype=np.uint32)
        # the whole in one call, into chain 62
        whole = an.merge_regions(bounds, owners, dst_chain=62)
        slab_whole = an.get_slab(62)
        # the same as two parts (even / odd regions), summed, into chain 63
        parts = []
        for parity in (0, 1):
            o = owners.copy()
            o[np.arange(nreg) % 2 != parity] = mg.NO_OWNER
            s = torch.zeros(n, dtype=torch.int64, device="cuda")
            a = torch.zeros(n, dtype=torch.int32, device="cuda")
            an.merge_export(bounds, o, s.data_ptr(), a.data_ptr())
            parts.append((s, a))
        s = parts[0][0] + parts[1][0]
        a = parts[0][1] + parts[1][1]
        torch.cuda.synchronize()
        halves = an.merge_import(s.data_ptr(), a.data_ptr(), dst_chain=63)
        assert halves == whole == port.slab_cost(data, slab_whole)
        got = an.get_slab(63)
        assert (got["type"] == slab_whole["type"]).all() and (got["dist"] == slab_whole["dist"]).all() and (got["len"] == slab_whole["len"]).all()
        an.close()
