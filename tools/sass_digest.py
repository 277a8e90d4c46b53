"""Per-kernel digests of a build's machine code:  python tools/sass_digest.py [lib.so] > profiles/rNN_committed_sass.md5
(`cuobjdump -sass`, one md5 per kernel over its non-empty lines; line information does not enter the listing).  The
committed digests are those of the build the last GPU tests of a round ran on; tests/test_abi.py rebuilds and compares."""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def digests(path: str) -> dict:
    out = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True, check=True).stdout
    result, name, lines = {}, None, []

    def flush():
        if name is not None:
            result[name] = hashlib.md5(("\n".join(lines) + "\n").encode()).hexdigest()

    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)$", line)
        if m:
            flush()
            name, lines = m.group(1), []
        if name is not None and line.strip():
            lines.append(line)
    flush()
    return result


if __name__ == "__main__":
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "megalania_b200", "_build", "libmegalania_cuda.so")
    for k, d in digests(lib).items():
        print(f"{d}  {k}")
