"""Turns an .ncu-rep (ncu --set full) into the short text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N] > profiles/rNN_name.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    path = sys.argv[1]
    nsrc = int(sys.argv[sys.argv.index("--source") + 1]) if "--source" in sys.argv else 0
    hdr, units, launches = raw(path)
    for row in launches:
        d = dict(zip(hdr, row))
        print(f"== {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in hdr:
            if k in WANT or any(k.startswith(w + ".") for w in ("dram__bytes_read.sum", "dram__bytes_write.sum")) and k.endswith("per_second"):
                print(f"  {k:72s} {d[k]:>22s} {units[hdr.index(k)]}")
        stalls = [(float(d[k]), k) for k in hdr if k.startswith(STALL_PREFIX) and k.endswith("_per_issue_active.ratio") and d[k]]
        for v, k in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {k[len(STALL_PREFIX):-len('_per_issue_active.ratio')]:40s} {v:8.3f} warps/issue")
    if nsrc:
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h = None
        body = []
        for r in rows:
            if len(r) > 4 and r[0] == "Address":
                h = r
                continue
            if h is not None and len(r) >= len(h) - 1 and r[0].startswith("0x"):
                body.append(r)
        if h:
            ci, si, ei = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
            total = sum(float(r[ci]) for r in body) or 1
            texec = sum(float(r[ei]) for r in body) or 1
            print(f"-- hottest {nsrc} SASS instructions by warp samples (of {int(total)}; {int(texec)} warp instructions executed)")
            order = sorted(range(len(body)), key=lambda i: -float(body[i][ci]))
            for i in order[:nsrc]:
                r = body[i]
                print(f"  {float(r[ci]) / total * 100:6.2f}% samples {float(r[ei]) / texec * 100:6.2f}% exec  #{i:5d} {r[si].strip()[:90]}")


if __name__ == "__main__":
    main()
