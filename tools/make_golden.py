"""Generates tests/golden/reference_vectors.json from the UNMODIFIED reference.

Run in the build container (needs /root/reference): it compiles the reference sources into
oracle/_ref (oracle/Makefile), drives them on deterministic inputs (tools/corpus.py) and
records known answers for the three parity functions plus an annealing trace, so that the
oracle port stays pinned where the reference tree is absent (the GPU box).

    python tools/make_golden.py
"""
from __future__ import annotations

import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle_lib as ol  # noqa: E402
from tools import corpus  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_vectors.json")
CASES = [("text", 2048, 12345), ("binary", 2048, 7), ("mixed", 4096, 7), ("text", 16384, 12345)]


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def packets_list(p):
    return [[int(x["type"]), int(x["dist"]), int(x["len"])] for x in p]


def packets_digest(p: np.ndarray) -> str:
    flat = np.stack([p["type"].astype(np.uint32), p["dist"].astype(np.uint32), p["len"].astype(np.uint32)], axis=-1)
    return sha(np.ascontiguousarray(flat).tobytes())


def boundaries(slab):
    out, p = [], 0
    while p < slab.size:
        out.append(p)
        p += int(slab[p]["len"])
    return out


def main() -> None:
    ol.build(force=False)
    ref = ol.Ref()
    port = ol.Port()
    doc = {"generator": "tools/make_golden.py", "reference": "blackle/Megalania @ /root/reference (unmodified, gcc -O3 -flto)",
           "sizeof": {"LZMAPacket": int(ref.lib.mgref_sizeof_packet()), "LZMAState": int(ref.lib.mgref_sizeof_state())},
           "price_table_sha256": sha(ref.price_table().astype("<u8").tobytes()),
           "price_table_samples": {str(i): int(ref.price_table()[i]) for i in (0, 1, 2, 31, 1024, 2017, 2047)},
           "rand_1673551_first16": [int(x) for x in ref.rand_stream(1673551, 16)],
           "heap": {}, "hello": {}, "cases": [], "cli": []}

    # max_heap tie semantics: streaming top-k with `<=` replacement (top_k_packet_finder.c:89)
    rng = corpus.SplitMix64(2024)
    keys = [int(rng.below(12)) for _ in range(200)]
    doc["heap"] = {"keys": keys, "k": 20, "pop_order": [int(x) for x in ref.heap_topk(keys, 20)]}

    # "hello hello" (the reference's own test string)
    hello = b"hello hello"
    lit = ol.literal_slab(11)
    s2 = lit.copy()
    s2[6] = (ol.MATCH, 5, 5)
    pops, counts = ref.topk_many(hello, lit, 0, np.arange(11))
    doc["hello"] = {
        "substring_counts": [ref.substring_count(hello, i) for i in range(11)],
        "substring_counts_max3": [ref.substring_count(hello, i, 3) for i in range(11)],
        "cost_literal": ref.slab_cost(hello, lit), "cost_match": ref.slab_cost(hello, s2),
        "bytes_match_hex": ref.encode_slab(hello, s2).hex(),
        "topk_mode0": [packets_list(pops[i][:counts[i]]) for i in range(11)],
    }

    for kind, n, seed in CASES:
        data = corpus.make(kind, n, seed)
        lit = ol.literal_slab(n)
        greedy = port.greedy_slab(data)  # deterministic slab builder; its digest is recorded below
        b = boundaries(greedy)
        case = {"kind": kind, "n": n, "seed": seed, "data_sha256": sha(data),
                "greedy_slab_digest": packets_digest(greedy), "greedy_live": len(b),
                "cost_literal": ref.slab_cost(data, lit), "cost_greedy": ref.slab_cost(data, greedy),
                "bytes_literal_sha256": sha(ref.encode_slab(data, lit)),
                "bytes_greedy_sha256": sha(ref.encode_slab(data, greedy)),
                "bytes_greedy_len": len(ref.encode_slab(data, greedy))}
        stop = b[len(b) // 2]
        m = ref.model_after_prefix(data, greedy, stop)
        case["model_mid"] = {"stop": int(stop), "sha256": sha(m.tobytes()), "cost": int(m["cost"]),
                             "ctx_state": int(m["ctx_state"]), "dists": [int(x) for x in m["dists"]]}
        if n <= 4096:
            pos = np.arange(n)
            pops, counts = ref.topk_many(data, lit, 0, pos)
            case["topk_mode0_digest"] = packets_digest(pops)
            case["topk_mode0_counts_sha256"] = sha(counts.astype("<i4").tobytes())
            pops1, counts1 = ref.topk_many(data, greedy, 1, b)
            case["topk_mode1_digest"] = packets_digest(pops1)
            case["topk_mode1_counts_sha256"] = sha(counts1.astype("<i4").tobytes())
            sample = [int(x) for x in np.linspace(1, n - 2, 6).astype(int)]
            case["topk_mode0_samples"] = {str(p): packets_list(pops[p][:counts[p]]) for p in sample}
        # annealing replay, glibc rand stream, seed of main.c:68
        for step in (0, 1):
            s = lit.copy() if step == 0 else greedy.copy()
            best = s.copy()
            attempts, bc, cc, trace = ref.anneal_epoch(data, s, best, 0, 0, step=step, evals=150)
            case[f"anneal_step{step}"] = {"evals": 150, "attempts": attempts, "best_cost": bc, "cur_cost": cc,
                                          "trace_sha256": sha(trace.tobytes()),
                                          "first_costs": [int(x) for x in trace["cost"][:12]],
                                          "slab_digest": packets_digest(s), "best_digest": packets_digest(best)}
        doc["cases"].append(case)

    # stock CLI end to end on tiny inputs (3 x 200 x n iterations; seconds)
    cli = os.path.join(ol.HERE, "_ref", "megalania_ref")
    for name, blob in (("hello", hello), ("text96", corpus.text(96, 5)), ("binary64", corpus.binary(64, 3))):
        with tempfile.NamedTemporaryFile(delete=False) as f:
            f.write(blob)
            path = f.name
        out = subprocess.run([cli, path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        os.unlink(path)
        doc["cli"].append({"name": name, "input_hex": blob.hex(), "output_hex": out.hex()})

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
