"""Per-simulation latency of the exact-mode stub kernel against the batch size (is it memory latency?)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
from blokus_self_play import SelfPlay, Config
cfg = Config(sims_per_move=800, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03,
             exploration_fraction=0.25, seed=1)
for n in (1, 16, 148, 296, 592, 1024, 1776):
    sp = SelfPlay(n, cfg)
    sp.run_stub(2); sp.reset()
    ms = sp.run_stub(-1)
    c = sp.counters()
    sims = c["sims"] - 2 * 800 * n
    plies = max(len(h) for h in sp.env.history())
    print(f"n={n:5d} kernel_ms={ms:8.2f} sims/s={sims/(ms*1e-3):.3e} longest_game_plies={plies} us_per_sim_of_longest={ms*1e3/(plies*800):.3f}", flush=True)
    sp.close()
