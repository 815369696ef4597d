"""Scratch timing probe: fused stub self-play kernel (config 3 shape) — sims/s from CUDA events."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
from blokus_self_play import SelfPlay, Config, Lib
LIB = Lib(os.environ['BK_LIB']) if os.environ.get('BK_LIB') else None
CASES = [(8192, 12)] if os.environ.get('BK_BIG') else [(1024, -1)] if os.environ.get('BK_FULLGAME') else [(1024, 16)] if os.environ.get('BK_QUICK') else [(1024, 4), (1024, 16), (4096, 8), (8192, 8)]
cfg = Config(sims_per_move=800, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03,
             exploration_fraction=0.25, seed=1)
if os.environ.get('BK_N'):
    CASES = [(int(os.environ['BK_N']), int(os.environ.get('BK_PLIES', '8')))]
for n, plies in CASES:
    sp = SelfPlay(n, cfg, lib=LIB, max_children_per_game=int(os.environ.get('BK_CAP', '0')))
    c0 = sp.counters()
    t = time.time()
    ms = sp.run_stub(plies)
    wall = time.time() - t
    c = sp.counters()
    sims = c["sims"] - c0["sims"]
    print(f"n={n} plies={plies} sims={sims} kernel_ms={ms:.2f} wall={wall*1e3:.1f} sims/s={sims/(ms*1e-3):.3e} "
          f"applies={c['applies']} movegens={c['movegens']} entries={c['entries']} nodes={c['nodes']}", flush=True)
    if plies < 0:
        sp.close()
        continue
    t = time.time()
    ms = sp.run_stub(plies)
    c2 = sp.counters()
    sims = c2["sims"] - c["sims"]
    print(f"   next {plies} plies: kernel_ms={ms:.2f} sims/s={sims/(ms*1e-3):.3e}", flush=True)
    sp.close()
