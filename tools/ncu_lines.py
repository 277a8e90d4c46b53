"""Executed warp instructions and stall samples per SOURCE LINE of an `ncu --set full --import-source on`
capture (kernels built with -lineinfo):  python tools/ncu_lines.py prof.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fpath, hdr, lines = "", None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if len(r) > 4 and r[0] == "Line No":
            hdr = r
            continue
        if hdr and r and r[0].isdigit():
            ie = hdr.index("Instructions Executed")
            sm = hdr.index("# Samples")
            def num(x):
                try:
                    return float(x)
                except ValueError:
                    return 0.0
            lines.append((fpath, int(r[0]), r[1].strip(), num(r[ie]), num(r[sm])))
    texec = sum(l[3] for l in lines) or 1
    tsamp = sum(l[4] for l in lines) or 1
    print(f"total warp instructions {texec:.4g}, samples {tsamp:.0f}")
    for f, ln, src, ex, sa in sorted(lines, key=lambda l: -l[3])[:top]:
        print(f"{100 * ex / texec:6.2f}% exec {100 * sa / tsamp:6.2f}% samples  {f}:{ln}  {src[:110]}")


if __name__ == "__main__":
    main()
