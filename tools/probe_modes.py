"""Row f3 probe: config-3 shape (1024 games, 800 sims/move, complete games) in the exact mode and with the
forced-ply shortcut: kernel time, simulations run, games/s."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
from blokus_self_play import SelfPlay, Config, MODE_SKIP_FORCED, MODE_TREE_REUSE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = Config(sims_per_move=800, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03,
             exploration_fraction=0.25, seed=1)
out = {}
for name, flags in (("exact", 0), ("skip_forced", MODE_SKIP_FORCED), ("tree_reuse", MODE_TREE_REUSE), ("tree_reuse+skip_forced", MODE_TREE_REUSE | MODE_SKIP_FORCED)):
    sp = SelfPlay(n, cfg)
    sp.set_mode(flags, 1)
    ms = sp.run_stub(-1)
    c = sp.counters()
    plies = sum(len(r) for r in sp.policy_records())
    out[name] = {"games": n, "kernel_ms": ms, "sims_run": c["sims"], "plies": plies, "plies_searched_equiv": c["sims"] / 800.0,
                 "games_per_s": n / (ms * 1e-3), "sims_per_s": c["sims"] / (ms * 1e-3)}
    sp.close()
out["forced_ply_fraction"] = 1.0 - out["skip_forced"]["plies_searched_equiv"] / out["exact"]["plies"]
out["tree_reuse_sims_saved"] = 1.0 - out["tree_reuse"]["sims_run"] / out["exact"]["sims_run"]
out["speedup_games_per_s"] = out["skip_forced"]["games_per_s"] / out["exact"]["games_per_s"]
print(json.dumps(out))
