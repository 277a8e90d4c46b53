"""Deterministic synthetic inputs for the BASELINE.json configs.

All randomness comes from splitmix64 so that any language can regenerate the same bytes.
  text(n, seed)    English-like: Zipf(1/rank) vocabulary of 5 000 pseudo-words, sentences
  binary(n, seed)  executable-like: opcode/modrm templates, zero runs, random islands
  mixed(n, seed)   first half text, second half binary (config 2 / 4)
  corpus16(n, seed) text + binary with long-range block repeats (config 5)
"""
from __future__ import annotations

import bisect

MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def below(self, n: int) -> int:
        return self.next() % n


_LETTERS = "etaoinshrdlcumwfgypbvkjxqz"
_LETTER_W = [127, 91, 82, 75, 70, 67, 63, 61, 60, 43, 40, 28, 28, 24, 24, 22, 20, 20, 19, 15, 10, 8, 2, 2, 1, 1]


def _cum(weights):
    out, acc = [], 0
    for w in weights:
        acc += w
        out.append(acc)
    return out


def _vocabulary(rng: SplitMix64, size: int = 5000):
    cum = _cum(_LETTER_W)
    words, seen = [], set()
    while len(words) < size:
        length = 2 + rng.below(8)
        w = "".join(_LETTERS[bisect.bisect_right(cum, rng.below(cum[-1]))] for _ in range(length))
        if w not in seen:
            seen.add(w)
            words.append(w)
    return words


def text(n: int, seed: int = 12345) -> bytes:
    rng = SplitMix64(seed)
    vocab = _vocabulary(rng)
    # Zipf 1/rank, integer weights
    cum = _cum([1_000_000 // (r + 1) for r in range(len(vocab))])
    out = bytearray()
    while len(out) < n:
        words = 5 + rng.below(16)
        for i in range(words):
            w = vocab[bisect.bisect_right(cum, rng.below(cum[-1]))]
            if i == 0:
                w = w.capitalize()
            out += w.encode()
            if i + 1 < words:
                out += b", " if rng.below(12) == 0 else b" "
        end = rng.below(10)
        out += b"?" if end == 0 else b"."
        out += b"\n" if rng.below(4) == 0 else b" "
    return bytes(out[:n])


def binary(n: int, seed: int = 7) -> bytes:
    rng = SplitMix64(seed)
    templates = []
    for _ in range(96):
        length = 2 + rng.below(5)
        templates.append(bytes(rng.below(256) for _ in range(length)))
    out = bytearray()
    while len(out) < n:
        kind = rng.below(10)
        if kind < 6:
            t = bytearray(templates[rng.below(len(templates))])
            if rng.below(3) == 0:
                t[-1] = rng.below(256)  # immediate / displacement byte
            out += t
        elif kind < 9:
            out += bytes(4 + rng.below(29))
        else:
            out += bytes(rng.below(256) for _ in range(4 + rng.below(61)))
    return bytes(out[:n])


def mixed(n: int, seed: int = 7) -> bytes:
    half = n // 2
    return text(half, seed) + binary(n - half, seed)


def corpus16(n: int, seed: int = 99) -> bytes:
    half = n // 2
    base = bytearray(text(half, seed) + binary(n - half, seed))
    rng = SplitMix64(seed ^ 0xC0FFEE)
    # re-insert blocks at long distances (1/16 .. 3/4 of the corpus)
    for _ in range(max(1, n // (256 * 1024))):
        blen = 256 + rng.below(4096)
        dist = n // 16 + rng.below(max(1, n * 3 // 4 - n // 16))
        dst = dist + rng.below(max(1, n - dist - blen))
        src = dst - dist
        if src >= 0 and dst + blen <= n:
            base[dst:dst + blen] = base[src:src + blen]
    return bytes(base)


GENERATORS = {"text": text, "binary": binary, "mixed": mixed, "corpus16": corpus16}


def make(kind: str, n: int, seed: int | None = None) -> bytes:
    fn = GENERATORS[kind]
    return fn(n) if seed is None else fn(n, seed)


if __name__ == "__main__":
    import argparse
    import sys

    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("kind", choices=sorted(GENERATORS))
    ap.add_argument("size", type=int)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("-o", "--output", default="-")
    a = ap.parse_args()
    blob = make(a.kind, a.size, a.seed)
    if a.output == "-":
        sys.stdout.buffer.write(blob)
    else:
        with open(a.output, "wb") as f:
            f.write(blob)
