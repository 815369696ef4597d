"""Small fixed workloads for `ncu --set full` captures (one GPU, a handful of launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
from blokus_self_play import GameBatch, SelfPlay, Config

what = sys.argv[1] if len(sys.argv) > 1 else "playout"
if what == "playout":
    b = GameBatch(4096)
    for k in range(4):
        b.reset()
        r = b.playout(seed=k)
    print("playout", r["total_steps"], r["kernel_ms"])
elif what == "conv":
    import torch
    from blokus_self_play.tc_resnet import to_padded_nhwc, conv3x3
    dev = torch.device("cuda", 0)
    xp = to_padded_nhwc(torch.relu(torch.randn(1024, 256, 20, 20, device=dev)))
    w9 = (torch.randn(9, 256, 256, device=dev) * 0.02).to(torch.bfloat16)
    bias = torch.zeros(256, device=dev)
    out = torch.empty_like(xp)
    for _ in range(4):
        conv3x3(xp, w9, bias, xp, True, 1024, out=out)
    torch.cuda.synchronize()
    print("conv done")
elif what == "mcts_big":
    # config-5 regime: 8192 games per GPU (the 20-games-per-SM instantiation), opening plies
    cfg = Config(sims_per_move=800, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03,
                 exploration_fraction=0.25, seed=1)
    sp = SelfPlay(8192, cfg, max_children_per_game=16384)
    for k in range(3):
        ms = sp.run_stub(2)
    print("mcts_big", sp.counters(), ms)
elif what == "mcts_mid":
    # mid-game MCTS: 120 plies unprofiled, then 2-ply launches (profile the LAST: ncu -k regex:k_selfplay_stub -s 3 -c 1)
    cfg = Config(sims_per_move=800, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03,
                 exploration_fraction=0.25, seed=1)
    sp = SelfPlay(1024, cfg, max_children_per_game=65536)
    sp.set_mode(1, 1)          # forced-ply shortcut only to reach ply 120 sooner (same games either way)
    sp.run_stub(60); sp.run_stub(60)
    sp.set_mode(0, 1)
    for k in range(2):
        ms = sp.run_stub(2)
    print("mcts_mid", sp.counters(), ms)
else:
    cfg = Config(sims_per_move=800, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03,
                 exploration_fraction=0.25, seed=1)
    sp = SelfPlay(1024, cfg, max_children_per_game=16384)
    for k in range(3):
        ms = sp.run_stub(2)
    print("mcts", sp.counters(), ms)
