"""Small workload touching every kernel of the library, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from blokus_self_play import (GameBatch, SelfPlay, Config, PLAYOUT_HASH, MODE_SKIP_FORCED, MODE_TREE_REUSE, host_evaluator)
import parity

b = GameBatch(6)
b.playout(seed=3, flags=PLAYOUT_HASH)
b.fetch(); b.scores(); b.payoff(); b.board(); b.board_state(); b.anchors(); b.legal_mask(); b.history()
b.reset()
for _ in range(12):
    lt = b.legal_tiles()
    b.apply([t[0] if t else -1 for t in lt])
c = b.clone(); c.playout(seed=1, max_plies=9); c.playout(seed=1)
cfg = Config(sims_per_move=24, sample_moves=3, c_base=19652, c_init=1.25, dirichlet_alpha=0.3, exploration_fraction=0.25, seed=2)
batched, _ = parity.fixed_network(1)
ev = host_evaluator(batched)
for flags, k in ((0, 1), (MODE_SKIP_FORCED, 1), (MODE_TREE_REUSE, 1), (MODE_TREE_REUSE | MODE_SKIP_FORCED, 4), (0, 5)):
    sp = SelfPlay(5, cfg)
    sp.set_mode(flags, k)
    if k == 1:
        sp.run_stub(7)
    sp.run_evaluator(ev, max_plies=6)
    sp.policy_records(); sp.policy_records_packed(); sp.last_root(); sp.training_tensors()
    sp.close()
print("sanitize target done")
