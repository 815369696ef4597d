"""BASELINE.json configs[3] as stated — MCTS with 800 sims/move and batched leaf evaluation on a random-init
ResNet(20,256) — played to the END of the games, the whole round inside the library (bk_selfplay_run_network: planes ->
41 tcgen05 convolutions + fused heads -> expand/backup, 8 bytes read back per round).

A full 1024-game run is ~2.2e8 evaluations = ~55 min of a B200 at the tensor roofline, so the complete-game measurement
uses fewer games with several leaves per round (virtual loss) such that a round still carries
~1024 positions: --games 64 --leaves 16 (default).  Prints one JSON line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
import numpy as np
import torch
from blokus_self_play import SelfPlay, Config
from blokus_self_play.resnet import ResNet
from blokus_self_play.tc_resnet import TensorCoreLeafEvaluator

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=64)
ap.add_argument("--leaves", type=int, default=16)
ap.add_argument("--sims", type=int, default=800)
ap.add_argument("--blocks", type=int, default=20)
ap.add_argument("--max-plies", type=int, default=-1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(1)
model = ResNet(a.blocks, 256).to(dev).eval()
ev = TensorCoreLeafEvaluator(model, max_rows=a.games * a.leaves)
cfg = Config(sims_per_move=a.sims, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03, exploration_fraction=0.25, seed=1)
sp = SelfPlay(a.games, cfg)
sp.set_mode(0, a.leaves)
warm = SelfPlay(a.games, cfg)
warm.set_mode(0, a.leaves)
warm.run_network(ev, max_plies=1)
warm.close()
torch.cuda.synchronize()
t0 = time.perf_counter()
info = sp.run_network(ev, max_plies=a.max_plies)
torch.cuda.synchronize()
secs = time.perf_counter() - t0
c = sp.counters()
hist = sp.env.history()
recs = sp.policy_records()
ok = all(int(v.sum()) == a.sims for r in recs for _, v in r) and all(len(r) == len(h) for r, h in zip(recs, hist))
flops_per_leaf = 2 * 400 * (5 * 9 * 256 + 2 * a.blocks * 256 * 256 * 9 + 2 * 256) + 2 * 400 * 4
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
print(json.dumps({
    "config": f"configs[3]: {a.games} games x {a.leaves} leaves per round, {a.sims} sims/move, ResNet({a.blocks},256) random init, eval mode, "
              f"native evaluator inside bk_selfplay_run_network; " + ("complete games" if a.max_plies < 0 else f"first {a.max_plies} plies"),
    "games": a.games, "finished_games": int(sp.env.is_terminal().sum()), "plies_searched": int(sum(len(h) for h in hist)),
    "sims": c["sims"], "evaluator_rounds": info["rounds"], "positions_evaluated": info["evals"],
    "mean_positions_per_round": info["evals"] / max(info["rounds"], 1), "seconds": secs, "device_ms": info["ms"],
    "sims_per_s": c["sims"] / secs, "positions_per_s": info["evals"] / secs, "games_per_s": a.games / secs,
    "tensor_tflops_useful": flops_per_leaf * info["evals"] / secs / 1e12,
    "frac_of_sustained_bf16_peak": flops_per_leaf * info["evals"] / secs / 1e12 / peaks.get("bf16_tflops_sustained", 1404.7),
    "invariants_ok": bool(ok), "every_policy_sums_to_sims": bool(ok)}), flush=True)
