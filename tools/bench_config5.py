"""BASELINE.json configs[4] per-GPU share: 8192 full self-play games on one B200, 800 sims/move, stub
evaluator, to completion (65 536 games = 8 such shards with global ids; no collective).  One JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
from blokus_self_play import SelfPlay, Config
from blokus_self_play.shard import shard_range
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "8"))
first, n = shard_range(65536, rank, world)
cfg = Config(sims_per_move=800, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03, exploration_fraction=0.25, seed=20261018)
t0 = time.time()
sp = SelfPlay(n, cfg, first_game_id=first, device=int(os.environ.get("LOCAL_RANK", "0")))
ms = sp.run_stub(-1)
c = sp.counters()
t1 = time.time()
st, po, va, offs = sp.training_tensors() if os.environ.get("BK_TENSORS") else (None, None, None, None)
print(json.dumps({"config": f"configs[4] shard: games {first}..{first + n - 1} of 65536 ({n} on this GPU), 800 sims/move, stub evaluator, to completion",
                  "games": n, "finished": int(sp.env.is_terminal().sum()), "sims": c["sims"], "kernel_ms": ms, "sims_per_s": c["sims"] / (ms * 1e-3),
                  "games_per_s": n / (ms * 1e-3), "moves_per_s": c["applies"] / (ms * 1e-3), "wall_s_incl_alloc": t1 - t0}))
