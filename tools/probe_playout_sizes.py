import sys, os
sys.path.insert(0, "/root/repo/blokus-engine_b200")
import numpy as np
from blokus_self_play import GameBatch
for n in (148, 592, 1184, 2368, 3552, 4096, 4736, 9472):
    b = GameBatch(n)
    best = 1e9
    for rep in range(4):
        b.reset()
        r = b.playout(seed=rep)
        best = min(best, r['kernel_ms'])
    st = r['steps']
    print(f"n={n} warps/SM={n/148:.1f} best_kernel_ms={best:.3f} moves/s={r['total_steps']/(best*1e-3):.3e} steps mean={st.mean():.1f} max={st.max()} min={st.min()} std={st.std():.1f}", flush=True)
    b.close()
