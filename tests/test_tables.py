"""The generated variant table (tools/gen_tables.py -> csrc/bk_tables_gen.h) against the oracle's
gen_variants and SURVEY.md Appendix B; and that the committed generated files are up to date."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_tables  # noqa: E402

COUNTS = [1, 2, 4, 2, 2, 8, 4, 1, 4, 2, 8, 8, 8, 4, 4, 4, 4, 8, 1, 8, 4]  # Appendix B


def test_generator_matches_oracle(orc):
    V = gen_tables.build()
    assert len(V) == 91
    k = 0
    for pid in range(21):
        assert orc.piece_num_variants(pid) == COUNTS[pid]
        for vi in range(COUNTS[pid]):
            ref = orc.piece_variant(pid, vi)
            v = V[k]
            assert (v["pid"], v["vi"]) == (pid, vi)
            assert v["offsets"] == ref["offsets"]
            assert v["w"] == ref["width"]
            assert (v["h"] - 1) * 20 + v["w"] == ref["len"]
            k += 1


def test_appendix_b_samples():
    V = {(v["pid"], v["vi"]): v for v in gen_tables.build()}
    assert V[(2, 1)]["offsets"] == [1, 20, 21]
    assert V[(5, 2)]["offsets"] == [1, 21, 40, 41] and V[(5, 2)]["w"] == 2
    assert V[(9, 1)]["offsets"] == [0, 20, 40, 60, 80]
    assert V[(16, 2)]["offsets"] == [0, 2, 20, 21, 22]
    assert V[(18, 0)]["offsets"] == [1, 20, 21, 22, 41]
    assert V[(19, 7)]["offsets"] == [1, 2, 20, 21, 41]
    assert V[(20, 3)]["offsets"] == [0, 20, 21, 22, 40]


def test_committed_generated_files_are_current():
    csrc = os.path.join(ROOT, "blokus-engine_b200", "csrc")
    with tempfile.TemporaryDirectory() as d:
        gen_tables.emit(d)
        for f in ("bk_tables_gen.h", "bk_movegen_gen.inc"):
            assert open(os.path.join(d, f)).read() == open(os.path.join(csrc, f)).read(), f


def test_library_tables_match_oracle(orc):
    """bk_piece_* entry points of the built sm_100a library (host tables; no GPU needed)."""
    import ctypes as C
    from blokus_self_play import Lib, DEFAULT_LIB
    lib = Lib(DEFAULT_LIB)
    for pid in range(21):
        assert lib.bk_piece_points(pid) == orc.piece_points(pid)
        assert lib.bk_piece_num_variants(pid) == orc.piece_num_variants(pid)
        for vi in range(orc.piece_num_variants(pid)):
            w, ln = C.c_int(), C.c_int()
            offs = (C.c_int * 5)()
            n = lib.bk_piece_variant(pid, vi, C.byref(w), C.byref(ln), offs)
            ref = orc.piece_variant(pid, vi)
            assert list(offs[:n]) == ref["offsets"] and w.value == ref["width"] and ln.value == ref["len"]
    assert lib.bk_piece_points(21) < 0 and b"bad piece id" in lib.bk_last_error()
