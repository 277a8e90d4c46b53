"""Shared-memory wavefront simulator for literal-run lane mappings (design aid, CPU only).

Replays the probability model over a plain-literal stream (the all-literal slab of the bench corpus) and counts,
for each candidate lane mapping, the shared-memory wavefronts the run loop would need: one wavefront serves at
most one 32-bit word per bank (same word = broadcast).  Used to choose the mapping in DESIGN.md section 4a.

  python tools/wavefront_sim.py [kind] [bytes]
"""
from __future__ import annotations

import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from tools import corpus  # noqa: E402

# slot map of mg_device.cuh (u16 slot indices)
S_LIT = 0
S_LIT01 = 768 + 40
S_ISMATCH = S_LIT01 + 22
WARP_STRIDE = 6608  # sizeof(WarpShared)
PROBS_BASE = 16384 + 1152 + 512  # offsetof(CtaShared, warp) (+ offsetof(Record, probs) = 0)
TRANS_BASE = 0


def lit_slots(b: int, row_rot: int = 0):
    """(slot, bit) for the nine events of a plain literal; row_rot rotates row r by r*row_rot words (layout variants)."""
    ev = [(S_ISMATCH, 0), (S_LIT01, b >> 7), (S_LIT01 + 4 + (b >> 7), (b >> 6) & 1)]
    row = b >> 6
    for d in range(2, 8):
        sub = (1 << (d - 2)) | ((b >> (8 - d)) & ((1 << (d - 2)) - 1))
        word = sub >> 1
        half = sub & 1
        word = (word + row * row_rot) % 32
        ev.append((S_LIT + row * 64 + word * 2 + half, (b >> (7 - d)) & 1))
    return ev


def wavefronts(addrs):
    """addrs: iterable of byte addresses accessed by one warp instruction (<= 4 bytes each)."""
    banks = {}
    for a in addrs:
        w = a >> 2
        banks.setdefault(w & 31, set()).add(w)
    return max((len(s) for s in banks.values()), default=0)


def wavefronts_wide(lane_addrs, width):
    """16- or 8-byte accesses: processed per quarter / half warp."""
    group = 8 if width == 16 else 16
    total = 0
    for g in range(0, 32, group):
        words = []
        for lane, a in lane_addrs:
            if g <= lane < g + group:
                words += [a + 4 * i for i in range(width // 4)]
        if words:
            total += wavefronts(words)
    return total


def update(p, bit):
    return p - (p >> 5) if bit else p + ((2048 - p) >> 5)


class Model:
    def __init__(self):
        self.p = {}

    def get(self, s):
        return self.p.get(s, 1024)

    def set(self, s, v):
        self.p[s] = v


def sim_current(data, warp=0):
    """One literal per step on lanes 0..8."""
    m = Model()
    base = PROBS_BASE + warp * WARP_STRIDE
    ev_base = base + 3728 + 256 + 128  # FindScratch.len_price
    wf = {"ev": 0, "ld": 0, "t": 0, "st": 0, "build": 0}
    n = len(data)
    for i, b in enumerate(data):
        ev = lit_slots(b)
        if i % 4 == 0:
            wf["ev"] += wavefronts_wide([(l, ev_base + 4 * 36 * l + 4 * (i % 32)) for l in range(9)], 16)
        if i % 32 == 0:
            wf["build"] += 9
        wf["ld"] += wavefronts([base + 2 * s for s, _ in ev])
        idx = [(m.get(s) | (bit << 11)) for s, bit in ev]
        wf["t"] += wavefronts([TRANS_BASE + 4 * x for x in idx])
        wf["st"] += wavefronts([base + 2 * s for s, _ in ev])
        for s, bit in ev:
            m.set(s, update(m.get(s), bit))
    return {k: v / n for k, v in wf.items()}


def sim_pairs(data, mode, row_rot=8, warp=0, ev_bytes=2):
    """Two consecutive literals per step: A on lanes 0..8, B on lanes 9..17.
    mode 't2'     : a B lane whose slot equals A's does ONE composite lookup (A's lane idles on the spare slot)
    mode 'chain'  : ... does two dependent lookups (second round only issues if any lane needs it)"""
    m = Model()
    base = PROBS_BASE + warp * WARP_STRIDE
    wf = {"ev": 0, "ld": 0, "t": 0, "t2": 0, "st": 0, "build": 0}
    n = len(data) & ~1
    dummy = base + 2 * 1838
    for i in range(0, n, 2):
        ea, eb = lit_slots(data[i], row_rot), lit_slots(data[i + 1], row_rot)
        if i % 32 == 0:
            wf["build"] += 9 * ev_bytes / 4
        per = 16 // ev_bytes  # events per 16-byte load
        if i % per == 0:
            wf["ev"] += 3  # 18 lanes x 16 B: three quarter-warps
        ld, t, t2, st = [], [], [], []
        for d in range(9):
            (sa, ba), (sb, bb) = ea[d], eb[d]
            pa = m.get(sa)
            if sa == sb:
                # A idles on the spare slot (probability 0, table entry 0), B does both
                ld += [dummy, base + 2 * sb]
                st += [dummy, base + 2 * sb]
                p1 = update(pa, ba)
                if mode == "t2":
                    t += [TRANS_BASE + 0, TRANS_BASE + 4 * (pa | ((4 | (ba << 1) | bb) << 11))]
                else:
                    t += [TRANS_BASE + 0, TRANS_BASE + 4 * (pa | (ba << 11))]
                    t2 += [TRANS_BASE + 4 * (p1 | (bb << 11))]
                m.set(sb, update(p1, bb))
            else:
                pb = m.get(sb)
                ld += [base + 2 * sa, base + 2 * sb]
                st += [base + 2 * sa, base + 2 * sb]
                t += [TRANS_BASE + 4 * (pa | (ba << 11)), TRANS_BASE + 4 * (pb | (bb << 11))]
                m.set(sa, update(pa, ba))
                m.set(sb, update(pb, bb))
        wf["ld"] += wavefronts(ld)
        wf["t"] += wavefronts(t)
        wf["t2"] += wavefronts(t2)
        wf["st"] += wavefronts(st)
    return {k: v / n for k, v in wf.items()}


def sim_cross(data, k, warp_stride=WARP_STRIDE):
    """k independent chains per warp instruction (lanes 9c..9c+8 = chain c), streams taken from k different
    places of the corpus, models k warp blocks apart."""
    n = len(data) // k
    ms = [Model() for _ in range(k)]
    wf = {"ev": 0, "ld": 0, "t": 0, "st": 0, "build": 0}
    for i in range(n):
        ld, t = [], []
        for c in range(k):
            base = PROBS_BASE + c * warp_stride
            ev = lit_slots(data[c * n + i])
            ld += [base + 2 * s for s, _ in ev]
            t += [TRANS_BASE + 4 * (ms[c].get(s) | (bit << 11)) for s, bit in ev]
            for s, bit in ev:
                ms[c].set(s, update(ms[c].get(s), bit))
        if i % 4 == 0:
            wf["ev"] += (9 * k + 7) // 8
        if i % 32 == 0:
            wf["build"] += 9 * k
        wf["ld"] += wavefronts(ld)
        wf["t"] += wavefronts(t)
        wf["st"] += wavefronts(ld)
    return {kk: v / (n * k) for kk, v in wf.items()}


def report(name, r):
    print(f"{name:32s} total {sum(r.values()):.3f}  " + "  ".join(f"{k} {v:.3f}" for k, v in r.items()))


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "mixed"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    blob = corpus.make(kind, size)
    # sample: 32 KiB of each half
    parts = {"text": blob[100000:100000 + 32768], "binary": blob[size // 2 + 100000: size // 2 + 100000 + 32768]}
    for nm, d in parts.items():
        print(f"--- {nm} ({len(d)} literals): shared-memory wavefronts per literal")
        report("current (1 literal / step)", sim_current(d))
        for rot in (0, 8):
            report(f"pairs, chained, row_rot {rot}", sim_pairs(d, "chain", rot))
            report(f"pairs, composite, row_rot {rot}", sim_pairs(d, "t2", rot))
        for k in (2, 3):
            report(f"{k} chains per step", sim_cross(d, k))
