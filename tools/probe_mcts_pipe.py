"""Where the two warps of the pipelined stub kernel wait (library built with -DBK_PIPE_STATS, BK_LIB points at it):
config 3 shape, complete games or a window of plies."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
from blokus_self_play import SelfPlay, Config, Lib
lib = Lib(os.environ["BK_LIB"]) if os.environ.get("BK_LIB") else None
cfg = Config(sims_per_move=800, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03, exploration_fraction=0.25, seed=1)
n = int(os.environ.get("BK_GAMES", "1024"))
sp = SelfPlay(n, cfg, lib=lib)


def raw():
    c = np.zeros(16, dtype=np.uint64)
    sp.lib.check(sp.lib.bk_selfplay_counters_raw(sp._h, c.ctypes.data_as(C.c_void_p)))
    return c.astype(np.float64)


def probe():
    c = np.zeros(32, dtype=np.uint64)
    sp.lib.check(sp.lib.bk_selfplay_probe_stats(sp._h, c.ctypes.data_as(C.c_void_p)))
    return c.astype(np.float64)


for label, plies in (("plies 0-15", 16), ("plies 16-79", 64), ("plies 80-143", 64), ("plies 144-207", 64), ("rest", -1)):
    c0, p0 = raw(), probe()
    ms = sp.run_stub(plies)
    d, q = raw() - c0, probe() - p0
    sims = d[0]
    if sims == 0:
        break
    print(f"{label}: sims={sims:.0f} kernel_ms={ms:.1f} sims/s={sims / (ms * 1e-3):.3e} cycles/sim per game={d[8] / sims:.0f} "
          f"A waits {d[6] / max(d[8], 1):.1%} B waits {d[7] / max(d[8], 1):.1%} applies/sim={d[1] / sims:.3f} movegens/sim={d[2] / sims:.3f}", flush=True)
    if d[14] > 0:
        print(f"      select {d[9] / d[14]:.0f} cycles each, {d[10] / d[14]:.2f} levels -> {d[9] / max(d[10], 1):.0f} cycles/level; voided selects {d[15] / d[14]:.2%}; "
              f"backup {d[11] / sims:.0f} cycles/sim; B: load+apply {d[12] / max(d[1], 1):.0f}, expand+link {d[13] / max(d[1], 1):.0f} cycles per leaf", flush=True)
    if q[9] > 0:
        print(f"      per level: root {q[8] / q[9]:.0f} cycles; other levels reading <= 32 entries {q[10] / max(q[11], 1):.0f} cycles ({q[11] / sims:.2f} per sim); "
              f"wider {q[12] / max(q[13], 1):.0f} cycles ({q[13] / sims:.2f} per sim); entries read per level {q[14] / max(q[9] + q[11] + q[13], 1):.1f}", flush=True)
