"""include/blokus_b200.hpp — the C++ host-side mirror of the reference's `Game` — compiles against the C ABI
and behaves: loud failure without a device here; the Appendix C min-tile game on a B200 (gpu-marked)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "blokus-engine_b200", "lib")


def build(tmp_path, libdir=LIBDIR, libname="blokus_b200"):
    exe = str(tmp_path / ("cpp_mirror_check_" + libname))
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp_mirror_check.cpp"),
                    "-L", libdir, "-l" + libname, "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


def test_cpp_mirror_logic_on_the_warp_emulator(tmp_path, emu_lib):
    """The same program linked against the tests' CPU warp-emulator build of the kernel sources: the C++ mirrors of
    `Game` and of the self_play client drive real games (314-ply min-tile game, stub self-play) without a GPU."""
    r = subprocess.run([build(tmp_path, os.path.join(ROOT, "tests", "warp_emu"), "blokus_emu")], capture_output=True, text=True)
    assert r.returncode == 0 and "device ok" in r.stdout, r.stdout + r.stderr


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    r = subprocess.run([build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no-device ok" in r.stdout or "device ok" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(tmp_path):
    r = subprocess.run([build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "device ok" in r.stdout, r.stdout + r.stderr
