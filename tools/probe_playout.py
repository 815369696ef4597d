"""Scratch timing probe: 4096-game playout kernel time (CUDA events inside the library)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
import numpy as np
from blokus_self_play import GameBatch, Lib
LIB = Lib(os.environ['BK_LIB']) if os.environ.get('BK_LIB') else None
for n in (4096, 16384, 65536):
    b = GameBatch(n, lib=LIB)
    for rep in range(4):
        b.reset()
        t = time.time()
        r = b.playout(seed=rep)
        wall = time.time() - t
        print(f"n={n} rep={rep} steps={r['total_steps']} kernel_ms={r['kernel_ms']:.3f} wall_ms={wall*1e3:.1f} "
              f"steps/s={r['total_steps']/(r['kernel_ms']*1e-3):.3e} movegens={r['movegens']} lane_ops={r['lane_ops']:.3e}", flush=True)
    b.close()
