"""Warp samples and executed instructions of an `ncu --set full --import-source on` capture of anneal_kernel, summed
by code region (finder / whole-window literal loop / rest of the walk / proposal state machine):
    python tools/ncu_regions.py prof.ncu-rep
The line ranges below follow mg_kernels.cuh as committed with the capture; adjust them when the file moves."""
import csv,io,subprocess,sys,collections
path=sys.argv[1]
out=subprocess.run(["ncu","-i",path,"--page","source","--csv","--print-source","cuda,sass"],stdout=subprocess.PIPE,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
fpath,hdr,lines="",None,[]
for r in rows:
    if len(r)>=2 and r[0]=="File Path": fpath=r[1].split("/")[-1]; continue
    if len(r)>4 and r[0]=="Line No": hdr=r; continue
    if hdr and r and r[0].isdigit():
        ie=hdr.index("Instructions Executed"); sm=hdr.index("# Samples")
        def num(x):
            try: return float(x)
            except: return 0.0
        lines.append((fpath,int(r[0]),num(r[ie]),num(r[sm])))
te=sum(l[2] for l in lines); ts=sum(l[3] for l in lines)
groups=collections.OrderedDict()
def grp(f,ln):
    if f=="mg_finder.cuh": return "finder"
    if f=="mg_kernels.cuh":
        if 476<=ln<=645: return "walk_windows"
        if 645<ln<=1000: return "walk(other)"
        if ln>1000: return "anneal state machine"
        return "kernels<476"
    if f=="mg_device.cuh": return "device.cuh"
    return f
for f,ln,e,s in lines:
    g=grp(f,ln); a=groups.setdefault(g,[0,0]); a[0]+=e; a[1]+=s
for g,(e,s) in sorted(groups.items(), key=lambda x:-x[1][1]): print(f"{g:28s} exec {100*e/te:6.2f}%  samples {100*s/ts:6.2f}%")
