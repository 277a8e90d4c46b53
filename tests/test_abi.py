"""The C-ABI library: builds for sm_100a, loads without a GPU, exports every symbol the header
declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "megalania_cuda.h")


def declared_functions():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^MG_API [^;(]*?\b(mg_\w+)\(", text, flags=re.M)))


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "megalania_cuda.h"\n#include "encoder_interface.h"\n'
                   "_Static_assert(sizeof(LZMAPacket) == 12, \"packet layout\");\n"
                   "int main(void) { EncoderInterface e; OutputInterface o; (void)e; (void)o; return 0; }\n")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                    "-o", str(tmp_path / "t.o")], check=True)


def test_library_exports_every_declared_symbol():
    import megalania_b200 as mg
    from megalania_b200 import api
    lib = mg.load_library()
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/megalania_cuda.h but not exported"
    assert set(api.EXPORTS) == set(names)


def test_library_is_sm100a_only():
    from megalania_b200 import build
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_build_is_the_machine_code_the_gpu_tests_ran_on():
    """The round's last GPU runs tested a build whose per-kernel SASS digests are committed
    (profiles/r02_committed_sass.md5, tools/sass_digest.py).  Experiment switches were removed from the source
    afterwards: the product build must still be that machine code, kernel by kernel."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sys.path.insert(0, ROOT)
    from megalania_b200 import build
    from tools import sass_digest
    want = {}
    for line in open(os.path.join(ROOT, "profiles", "r02_committed_sass.md5")):
        digest, name = line.split()
        want[name] = digest
    got = sass_digest.digests(build.build_library())
    assert set(got) == set(want)
    changed = sorted(k for k in want if got[k] != want[k])
    assert not changed, f"kernels rebuilt to different machine code than the GPU-tested build: {changed}"


def test_no_cpu_fallback():
    import megalania_b200 as mg
    lib = mg.load_library()
    count = ctypes.c_int(0)
    try:
        rt = ctypes.CDLL("libcudart.so")
        have_gpu = rt.cudaGetDeviceCount(ctypes.byref(count)) == 0 and count.value > 0
    except OSError:
        have_gpu = os.path.exists("/dev/nvidia0")
    if have_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(mg.MegalaniaError) as err:
        mg.Context(b"hello hello")
    assert err.value.code == -2
    assert lib.mg_last_error()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "megalania_b200")
    for base, _, files in os.walk(pkg):
        if "_build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                text = open(os.path.join(base, f)).read()
                assert "oracle_lib" not in text and "mg_oracle" not in text and "libmegalania_ref" not in text, f


def test_cli_usage_and_error_exit_status(tmp_path):
    """The drop-in CLI keeps the reference's contract (src/main.c:29-38): one positional filename, progress and
    errors on stderr, exit status 255 on usage or I/O errors - and, here, when no CUDA device is present
    (there is no CPU fallback)."""
    import subprocess
    from megalania_b200 import build
    cli = build.build_cli() or build.CLI
    assert os.path.exists(cli)
    r = subprocess.run([cli], capture_output=True)
    assert r.returncode == 255 and b"usage:" in r.stderr and r.stdout == b""
    r = subprocess.run([cli, "--rounds"], capture_output=True)
    assert r.returncode == 255 and b"usage:" in r.stderr
    r = subprocess.run([cli, "--no-such-flag", "x"], capture_output=True)
    assert r.returncode == 255 and b"usage:" in r.stderr
    r = subprocess.run([cli, str(tmp_path / "missing.bin")], capture_output=True)
    assert r.returncode == 255 and r.stdout == b""
    import torch
    if not torch.cuda.is_available():
        f = tmp_path / "in.bin"
        f.write_bytes(b"hello hello hello")
        r = subprocess.run([cli, "--time", "1", str(f)], capture_output=True)
        assert r.returncode == 255 and r.stdout == b"" and b"mg_ctx_create" in r.stderr
