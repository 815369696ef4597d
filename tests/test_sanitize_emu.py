"""The kernel sources under AddressSanitizer + UndefinedBehaviorSanitizer on the CPU warp emulator (compute-sanitizer
is closed on the GPU pool): every kernel of the library runs a small workload; any out-of-bounds access to a pool,
shared array or host buffer, or undefined shift, fails the test.  tests/sanitize/README.md has the ThreadSanitizer
recipe (lane-to-lane races; slower, run by hand)."""
import glob
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find(name):
    for pat in ("/usr/lib/x86_64-linux-gnu/%s.so.*", "/usr/lib/gcc/x86_64-linux-gnu/*/%s.so", "/usr/lib64/%s.so.*"):
        hits = sorted(glob.glob(pat % name))
        if hits:
            return hits[0]
    return None


def test_kernel_sources_clean_under_asan_ubsan(tmp_path):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    asan, ubsan = _find("libasan"), _find("libubsan")
    if not (gxx and asan and ubsan):
        pytest.skip("no sanitizer runtimes in this image")
    emu = os.path.join(ROOT, "tests", "warp_emu")
    lib = str(tmp_path / "libblokus_emu_asan.so")
    srcs = []
    for s in sorted(glob.glob(os.path.join(ROOT, "blokus-engine_b200", "csrc", "*.cu"))):
        srcs += ["-x", "c++", s]
    subprocess.run([gxx, "-O1", "-g", "-std=c++20", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-w",
                    "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I", emu, "-shared", "-o", lib,
                    os.path.join(emu, "emu_runtime.cpp")] + srcs, check=True)
    env = dict(os.environ, LD_PRELOAD=f"{asan} {ubsan}", ASAN_OPTIONS="detect_leaks=0:halt_on_error=1",
               UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    r = subprocess.run(["python", os.path.join(ROOT, "tests", "sanitize", "target.py"), lib, "quick"], env=env,
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "sanitize target done" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-4000:]
