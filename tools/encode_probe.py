"""Range-coder pass timing against the reference's CPU pass: python tools/encode_probe.py [n] [kind]"""
import lzma
import sys
import time
sys.path.insert(0, '.')
import megalania_b200 as mg
from oracle import oracle_lib
from tools import corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
kind = sys.argv[2] if len(sys.argv) > 2 else "mixed"
data = corpus.make(kind, n)
ctx = mg.Context(data)
port = oracle_lib.Port()
cpu = oracle_lib.Ref() if oracle_lib.ref_available() else port
slabs = [("all-literal", mg.literal_slab(n))]
if n <= (4 << 20):
    slabs.append(("greedy", port.greedy_slab(data)))
for name, slab in slabs:
    ctx.encode_slab_buffer(slab)
    t = time.time(); out = ctx.encode_slab_buffer(slab); dt = time.time() - t
    st = ctx.encode_stats()
    t = time.time(); want = cpu.encode_slab(data, slab); cpu_dt = time.time() - t
    assert out == want
    assert lzma.decompress(out, format=lzma.FORMAT_ALONE) == data
    print(f"{name}: {len(out)} bytes, {st['events']} events, device {st['kernel_ms']:.2f} ms "
          f"({st['kernel_ms'] * 1e6 / max(1, st['events']) * 1.965:.1f} cycles/event at 1965 MHz), call {dt:.3f} s; "
          f"CPU ({type(cpu).__name__}) {cpu_dt:.3f} s")
