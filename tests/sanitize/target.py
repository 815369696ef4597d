"""TEST HARNESS ONLY.  A small workload that drives every kernel of the library through the CPU warp emulator
(tests/warp_emu): run under AddressSanitizer / UndefinedBehaviorSanitizer / ThreadSanitizer builds of the emulator
library it checks the KERNEL SOURCES for out-of-bounds accesses, undefined shifts and lane-to-lane data races
(every lane is an OS thread there, warp collectives are the only synchronisation, so a missing __syncwarp is a
reported race).  compute-sanitizer is not available on the GPU pool; this is the substitute.

    python tests/sanitize/target.py <path to an emulator build of the library>
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "blokus-engine_b200"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
from blokus_self_play import (Lib, GameBatch, SelfPlay, Config, PLAYOUT_HASH, MODE_SKIP_FORCED, MODE_TREE_REUSE)
import parity

lib = Lib(sys.argv[1])
quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
b = GameBatch(2 if quick else 3, lib=lib)
b.playout(seed=3, flags=PLAYOUT_HASH, max_plies=60 if quick else -1)
b.fetch(); b.scores(); b.payoff(); b.board(); b.board_state(); b.anchors(); b.legal_mask(); b.history()
b.reset()
for _ in range(12):
    b.apply([t[0] if t else -1 for t in b.legal_tiles()])
c = b.clone(); c.playout(seed=1, max_plies=9); c.playout(seed=1, max_plies=40 if quick else -1)
cfg = Config(sims_per_move=12 if quick else 16, sample_moves=3, c_base=19652, c_init=1.25, dirichlet_alpha=0.3,
             exploration_fraction=0.25, seed=2)
batched, _ = parity.fixed_network(1)
MODES = ((0, 1), (MODE_SKIP_FORCED, 1), (MODE_TREE_REUSE, 1), (MODE_TREE_REUSE | MODE_SKIP_FORCED, 4), (0, 5))
if quick:
    MODES = ((0, 1), (MODE_TREE_REUSE, 1), (MODE_TREE_REUSE | MODE_SKIP_FORCED, 4))
for flags, k in MODES:
    sp = SelfPlay(2 if quick else 3, cfg, lib=lib)
    sp.set_mode(flags, k)
    if k == 1:
        sp.run_stub(5 if quick else 7)
    sp.run_evaluator(batched, max_plies=4 if quick else 6, buffers=parity.HostBuffers())
    sp.policy_records(); sp.policy_records_packed(); sp.last_root(); sp.training_tensors(buffers=parity.HostBuffers())
    sp.close()
# the two-warp pipelined stub kernel (the emulator build only runs it on request)
os.environ["BK_STUB_PIPE"] = "1"
sp = SelfPlay(1, cfg, first_game_id=11, lib=lib)
del os.environ["BK_STUB_PIPE"]
sp.run_stub(3 if quick else 5)
sp.policy_records(); sp.last_root()
sp.close()
print("sanitize target done")
