"""Range-coder pass timing: python tools/encode_probe.py [n] [kind]"""
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
for name, slab in (("all-literal", mg.literal_slab(n)), ("greedy", oracle_lib.Port().greedy_slab(data))):
    ctx.encode_slab_buffer(slab)
    t = time.time(); out = ctx.encode_slab_buffer(slab); dt = time.time() - t
    assert lzma.decompress(out, format=lzma.FORMAT_ALONE) == data
    print(name, len(out), "bytes in %.3f s" % dt)
