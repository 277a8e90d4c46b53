"""The NCCL collectives behind the C ABI (mg_comm_*), one process per GPU.  Needs two GPUs (`gpurun --gpus 2`);
skipped on a single-GPU box.  The rule the ranks apply (mg_temper_decide) is covered on the CPU in test_host_logic."""
import multiprocessing as mp
import traceback

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 65536
CHAINS = 64


def _same(a, b):
    return (a["type"] == b["type"]).all() and (a["dist"] == b["dist"]).all() and (a["len"] == b["len"]).all()


def _rank_main(rank, world, uid, q_out, q_in):
    try:
        import megalania_b200 as mg
        from megalania_b200 import tempering
        from tools import corpus
        data = corpus.make("mixed", N)
        ctx = mg.Context(data, device=rank)
        ctx.comm_init(rank, world, uid)
        an = mg.Annealer(ctx, CHAINS, seed=100 + rank, checkpoint_stride=1024)
        an.set_slab(None)
        an.run(25)
        out = {}
        # ---- best-slab broadcast ----
        cur0, best0 = an.costs()
        winner, cost, chain = an.comm_exchange_best()
        cur1, best1 = an.costs()
        out["winner"], out["cost"] = winner, cost
        out["local_best"] = int(best0[best0 > 0].min())
        slab = an.get_slab(chain, best=(rank == winner))
        out["slab"] = slab
        out["slab_cost"] = ctx.score_slab(slab)
        if rank != winner:
            assert int(cur1[chain]) == cost
            assert chain == int(cur0.argmax())
        out["stats"] = ctx.comm_stats()
        # the installed chain keeps annealing from the copied checkpoints: its costs must stay exact
        an.run(15, first_eval=mg.CONTINUE_EVALS)
        cur2, _ = an.costs()
        out["after_install_exact"] = all(ctx.score_slab(an.get_slab(c)) == int(cur2[c]) for c in (chain, 0, CHAINS - 1))
        # ---- replica exchange ----
        ladder = tempering.temperature_ladder(world * CHAINS, 16.0, 65536.0)
        temps = ladder[rank::world].copy()
        new = an.comm_temper_exchange(temps, 3, seed=99)
        out["temps_in"], out["temps_out"], out["costs"] = temps, new, cur2.copy()
        # ---- any chain from any rank ----
        an.comm_broadcast_chain(world - 1, 3, 5)
        cur3, _ = an.costs()
        out["bc_slab"] = an.get_slab(3 if rank == world - 1 else 5)
        out["bc_cost"] = int(cur3[3 if rank == world - 1 else 5])
        an.run(10, first_eval=mg.CONTINUE_EVALS)
        cur4, _ = an.costs()
        out["after_broadcast_exact"] = ctx.score_slab(an.get_slab(5)) == int(cur4[5])
        # ---- region merge over the ranks ----
        nreg = 8
        bounds = np.linspace(0, N, nreg + 1).astype(np.uint32)
        owners = np.full(nreg, mg.api.NO_OWNER, dtype=np.uint32)
        mine = np.arange(rank, nreg, world)
        owners[mine] = (mine % CHAINS).astype(np.uint32)
        merged = an.comm_merge_regions(bounds, owners, dst_chain=7)
        cur5, _ = an.costs()
        out["merged"] = merged
        out["merged_slab"] = an.get_slab(7)
        out["merged_exact"] = ctx.score_slab(out["merged_slab"]) == merged == int(cur5[7])
        out["gathered"] = ctx.comm_allgather(1000 + rank)
        an.close()
        ctx.close()
        q_out.put((rank, out))
    except Exception:
        q_out.put((rank, traceback.format_exc()))


def test_collectives_two_ranks(port):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import megalania_b200 as mg
    from megalania_b200 import tempering
    mg.load_library()
    world = 2
    uid = mg.Context.comm_unique_id()
    ctx = mp.get_context("spawn")
    q_out, q_in = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, uid, q_out, q_in)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, out = q_out.get(timeout=600)
        assert isinstance(out, dict), out
        res[rank] = out
    for p in procs:
        p.join(timeout=60)
    a, b = res[0], res[1]
    from tools import corpus
    data = corpus.make("mixed", N)
    # best-slab broadcast: same verdict everywhere, the slab arrived intact and priced exactly
    assert a["winner"] == b["winner"] and a["cost"] == b["cost"]
    assert a["cost"] == min(a["local_best"], b["local_best"])
    assert _same(a["slab"], b["slab"])
    assert a["slab_cost"] == b["slab_cost"] == a["cost"] == port.slab_cost(data, a["slab"])
    loser = res[1 - a["winner"]]
    assert loser["stats"]["installs_by_copy"] + loser["stats"]["installs_by_rescore"] == 1
    assert a["after_install_exact"] and b["after_install_exact"]
    # replica exchange: every rank applied the global rule
    costs = np.concatenate([a["costs"], b["costs"]]).astype(np.int64)
    temps = np.concatenate([a["temps_in"], b["temps_in"]])
    want = tempering.exchange_temperatures(costs, temps, 3, seed=99)
    assert (np.concatenate([a["temps_out"], b["temps_out"]]) == want).all()
    # chain broadcast
    assert _same(a["bc_slab"], b["bc_slab"]) and a["bc_cost"] == b["bc_cost"] == port.slab_cost(data, a["bc_slab"])
    assert a["after_broadcast_exact"] and b["after_broadcast_exact"]
    # merge
    assert a["merged"] == b["merged"] and _same(a["merged_slab"], b["merged_slab"])
    assert a["merged_exact"] and b["merged_exact"]
    assert port.slab_valid(data, a["merged_slab"])
    assert list(a["gathered"]) == [1000, 1001] == list(b["gathered"])
