"""GPU tests at BASELINE.json's full sizes (1 MiB mixed corpus) and of the paths the small parity
tests do not reach: time-boxed steps, the best-slab journal overflow, replica exchange helpers,
the temperature schedule, the C host CLI."""
import lzma
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MIB = 1 << 20


def same_packets(a, b):
    return (a["type"] == b["type"]).all() and (a["dist"] == b["dist"]).all() and (a["len"] == b["len"]).all()


def boundaries(slab):
    out, p = [], 0
    while p < slab.size:
        out.append(p)
        p += int(slab[p]["len"])
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
def big_greedy(port, big):
    return port.greedy_slab(big)


@pytest.fixture(scope="module")
def big_ctx(mg, big):
    ctx = mg.Context(big)
    yield ctx
    ctx.close()


def test_1mib_cost_and_bytes(mg, port, big, big_ctx, big_greedy):
    lit = mg.literal_slab(MIB)
    assert big_ctx.score_slab(lit) == port.slab_cost(big, lit)
    greedy = big_greedy
    assert big_ctx.score_slab(greedy) == port.slab_cost(big, greedy)
    stream = big_ctx.encode_slab(greedy)
    assert stream == port.encode_slab(big, greedy)
    assert lzma.decompress(stream, format=lzma.FORMAT_ALONE) == big


def test_1mib_topk_sampled_positions(mg, port, big, big_ctx, big_greedy):
    """every position of the first 64 KiB under the init state is covered by
    the 64 KiB cases; here 2048 uniformly sampled positions of the full 1 MiB, both state modes."""
    rng = np.random.default_rng(7)
    lit = mg.literal_slab(MIB)
    pos = np.sort(rng.choice(MIB, 2048, replace=False)).astype(np.uint64)
    pops, prices, counts = big_ctx.find_topk(lit, pos, state_mode=0)
    wp, wprice, wc = port.topk_many_priced(big, lit, 0, pos)
    assert (counts == wc).all() and same_packets(pops, wp) and (prices == wprice).all()
    greedy = big_greedy
    b = boundaries(greedy)
    sample = np.sort(rng.choice(b, 1024, replace=False))
    pops, prices, counts = big_ctx.find_topk(greedy, sample, state_mode=1)
    wp, wprice, wc = port.topk_many_priced(big, greedy, 1, sample)
    assert (counts == wc).all() and same_packets(pops, wp) and (prices == wprice).all()


def test_1mib_anneal_trace(mg, port, big, big_ctx):
    chains, evals, seed = 2, 6, 4242
    an = mg.Annealer(big_ctx, chains, trace_capacity=1024, seed=seed)
    an.set_slab(None)
    st = an.run(evals)
    cur, best = an.costs()
    assert st["evals"] == chains * evals and st["log_overflows"] == 0
    lit = mg.literal_slab(MIB)
    for c in range(chains):
        slab, bslab = lit.copy(), lit.copy()
        attempts, bc, cc, _, trace = port.anneal_epoch(big, slab, bslab, 0, 0, rng_mode=1,
                                                       rng_state=port.chain_seed(seed, c), evals=evals)
        got = an.trace(c)
        assert len(got) == attempts and (got["cost"] == trace["cost"]).all() and (got["flags"] == trace["flags"]).all()
        assert int(cur[c]) == cc and int(best[c]) == bc
        assert same_packets(an.get_slab(c, best=True), bslab)
    an.close()


def test_long_run_journal_overflow_and_budget(mg, port, corpora):
    """Thousands of accepted edits overflow the best-slab journal (full-copy fallback); the same run is
    split into time-boxed steps on a second annealer and must land on the identical state."""
    n = 2048
    data = corpora("text", n)
    seed, evals = 31337, 3000
    lit = mg.literal_slab(n)
    with mg.Context(data) as ctx:
        a = mg.Annealer(ctx, 2, seed=seed, checkpoint_stride=512)
        a.set_slab(None)
        a.run(evals)
        b = mg.Annealer(ctx, 2, seed=seed, checkpoint_stride=512)
        b.set_slab(None)
        done = 0
        while done < 2 * evals:
            st = b.run(50, first_eval=mg.CONTINUE_EVALS, packet_budget=20000)
            done += st["evals"]
            if st["evals"] == 0:
                break
        for c in range(2):
            slab, bslab = lit.copy(), lit.copy()
            _, bc, cc, _, _ = port.anneal_epoch(data, slab, bslab, 0, 0, rng_mode=1, rng_state=port.chain_seed(seed, c),
                                                evals=evals)
            cur, best = a.costs()
            assert int(cur[c]) == cc and int(best[c]) == bc
            assert same_packets(a.get_slab(c), slab) and same_packets(a.get_slab(c, best=True), bslab)
            assert ctx.score_slab(a.get_slab(c, best=True)) == bc
        # the time-boxed annealer ran a different number of evaluations per chain; its invariants:
        cur, best = b.costs()
        for c in range(2):
            assert ctx.score_slab(b.get_slab(c)) == int(cur[c])
            assert ctx.score_slab(b.get_slab(c, best=True)) == int(best[c])
            assert port.slab_valid(data, b.get_slab(c, best=True))
        a.close()
        b.close()


def test_temperature_schedule_and_exchange_helpers(mg, port, corpora):
    n = 4096
    data = corpora("mixed", n)
    with mg.Context(data) as ctx:
        an = mg.Annealer(ctx, 8, seed=5)
        an.set_slab(None)
        temps = np.geomspace(50.0, 50000.0, 8).astype(np.float32)
        st = an.run(300, schedule=mg.SCHEDULE_TEMPERATURE, temperatures=temps)
        assert st["evals"] == 8 * 300
        cur, best = an.costs()
        assert (best <= cur).all() and (best > 0).all()
        for c in range(8):
            slab = an.get_slab(c)
            assert port.slab_valid(data, slab)
            assert port.slab_cost(data, slab) == int(cur[c])
            assert port.slab_cost(data, an.get_slab(c, best=True)) == int(best[c])
        # hotter chains accept more: coldest chain ends no worse than the hottest
        assert cur[0] <= cur[7]
        # swap two chains' slabs (replica exchange on one device) and keep annealing
        s0, s7 = an.get_slab(0), an.get_slab(7)
        an.swap_chains(0, 7)
        cur2, _ = an.costs()
        assert int(cur2[0]) == int(cur[7]) and int(cur2[7]) == int(cur[0])
        assert same_packets(an.get_slab(0), s7) and same_packets(an.get_slab(7), s0)
        an.run(50, schedule=mg.SCHEDULE_TEMPERATURE, temperatures=temps)
        cur3, best3 = an.costs()
        for c in (0, 7):
            assert port.slab_cost(data, an.get_slab(c)) == int(cur3[c])
            assert port.slab_cost(data, an.get_slab(c, best=True)) == int(best3[c])
        # export / import through a device buffer (what the NCCL broadcast uses)
        import torch
        buf = torch.empty(n * 8, dtype=torch.uint8, device="cuda:0")
        an.export_slab(3, True, buf.data_ptr())
        an.import_slab(5, buf.data_ptr(), adopt_cost=True)
        cur4, best4 = an.costs()
        assert int(cur4[5]) == int(best3[3])
        assert same_packets(an.get_slab(5), an.get_slab(3, best=True))
        an.close()


def test_oneshot_and_encode_roundtrip(mg, corpora):
    n = 65536
    data = corpora("text", n)
    with mg.Context(data) as ctx:
        best, cost, st = mg.anneal_oneshot(ctx, chains=64, evals=200, seed=3)
        assert st["evals"] == 64 * 200
        assert ctx.score_slab(best) == cost
        stream = ctx.encode_slab(best)
        assert lzma.decompress(stream, format=lzma.FORMAT_ALONE) == data
        # the reference's own size estimate (src/main.c:98); table prices are floors, so it runs slightly low
        assert abs(len(stream) - (18 + cost / 16384)) < 0.002 * len(stream)


def test_c_host_cli(tmp_path, corpora):
    from megalania_b200 import build
    cli = build.build_cli()
    assert cli and os.path.exists(cli)
    data = corpora("binary", 4096)
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    out = subprocess.run([cli, "--chains", "256", "--iters", "300", "--epochs", "1", str(src)], stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, check=True)
    assert lzma.decompress(out.stdout, format=lzma.FORMAT_ALONE) == data
    assert len(out.stdout) < len(lzma.compress(data, format=lzma.FORMAT_ALONE, preset=9 | lzma.PRESET_EXTREME)) + 400
    assert b"current file size" in out.stderr
    # usage / IO errors keep the reference's exit status (255)
    assert subprocess.run([cli]).returncode == 255
    assert subprocess.run([cli, str(tmp_path / "missing")], stderr=subprocess.DEVNULL).returncode == 255


def test_header_dictionary_covers_inputs_above_4mib(mg, corpora):
    """src/lzma_header_encoder.c:16 always writes 4 MiB; above 4 MiB of input that makes far matches
    undecodable, so the dictionary size grows to the next power of two (identical bytes up to 4 MiB)."""
    import lzma
    n = (4 << 20) + 4096
    data = corpora("corpus16", n)
    with mg.Context(data) as ctx:
        slab = mg.literal_slab(n)
        # one match that reaches farther than 4 MiB: the last 64 bytes copy an earlier block
        src = 100
        tail = n - 64
        patched = bytearray(data)
        patched[tail:] = patched[src:src + 64]
    data2 = bytes(patched)
    with mg.Context(data2) as ctx:
        slab = mg.literal_slab(n)
        slab[tail] = (mg.MATCH, tail - src - 1, 64)
        stream = ctx.encode_slab_buffer(slab)
        assert int.from_bytes(stream[1:5], "little") == 8 << 20
        assert int.from_bytes(stream[5:13], "little") == n
        assert lzma.decompress(stream, format=lzma.FORMAT_ALONE) == data2
    small = corpora("text", 4096)
    with mg.Context(small) as ctx:
        stream = ctx.encode_slab_buffer(mg.literal_slab(4096))
        assert int.from_bytes(stream[1:5], "little") == 0x400000
