"""BASELINE.json configs[3]: 1024 games, MCTS 800 sims/move, batched leaf evaluation on a random-init
ResNet(blocks=20, width=256) (bf16 autocast, channels_last, .eval()), one leaf per live game per round.
A whole game needs ~2.2e8 evaluations (~1 h at the tensor-core ceiling), so a fixed number of lockstep
evaluator rounds of the first ply is timed.  Prints one JSON line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
import torch
from blokus_self_play import SelfPlay, Config
from blokus_self_play.resnet import ResNet, LeafEvaluator

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=1024)
ap.add_argument("--rounds", type=int, default=40)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--blocks", type=int, default=20)
ap.add_argument("--width", type=int, default=256)
ap.add_argument("--fp32", action="store_true")
ap.add_argument("--leaves", type=int, default=1, help="row f3: leaves per game per evaluator round (virtual loss); 1 = exact mode")
ap.add_argument("--tc", action="store_true", help="hand-written tcgen05 trunk (bk_conv3x3_bf16) instead of cuDNN")
a = ap.parse_args()
torch.manual_seed(20261018)
dev = torch.device("cuda", 0)
model = ResNet(a.blocks, a.width).to(dev)
if a.tc:
    from blokus_self_play.tc_resnet import TensorCoreLeafEvaluator
    ev = TensorCoreLeafEvaluator(model)
else:
    ev = LeafEvaluator(model, bf16=not a.fp32)
cfg = Config(sims_per_move=800, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03, exploration_fraction=0.25, seed=1)
sp = SelfPlay(a.games, cfg)
sp.set_stream(torch.cuda.current_stream(dev).cuda_stream)
if a.leaves > 1:
    sp.set_mode(0, a.leaves)
planes = torch.zeros((a.games * a.leaves, 5, 20, 20), dtype=torch.float32, device=dev)
sp.begin_ply()
sp.leaf_planes(planes.data_ptr(), want_count=False)
t_eval = t_tree = 0.0
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
evals = 0
dense = a.leaves > 1
for r in range(a.warmup + a.rounds):
    rows = sp.leaf_rows()                              # the evaluator batch is dense (no finished games, no empty slots)
    e0.record()
    policy, value = ev(planes[:rows])
    policy, value = policy.contiguous(), value.contiguous()
    e1.record()
    sp.expand_backup(policy.data_ptr(), value.data_ptr(), want_count=False)
    sp.leaf_planes(planes.data_ptr(), want_count=False)
    e2.record()
    torch.cuda.synchronize()
    if r == a.warmup - 1:
        sims0 = sp.counters()["sims"]
    if r >= a.warmup:
        t_eval += e0.elapsed_time(e1)
        t_tree += e1.elapsed_time(e2)
        evals += rows
sims = sp.counters()["sims"] - sims0     # simulations actually completed (slots left empty by collisions do not count)
flops_per_leaf = 2 * 400 * (5 * 9 * a.width + 2 * a.blocks * a.width * a.width * 9 + 2 * a.width) + 2 * 400 * 4
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0}
tot = (t_eval + t_tree) * 1e-3
print(json.dumps({
    "metric": "mcts_leaf_evals_per_sec_resnet", "value": sims / tot, "unit": "sims/s",
    "leaves_per_round": a.leaves, "batch_slots_per_round": a.games * a.leaves, "slot_fill": sims / max(evals, 1),
    "config": f"configs[3]: {a.games} games, 800 sims/move, ResNet({a.blocks},{a.width}) random init, {'hand-written tcgen05 convolutions (bf16 operands, f32 accumulate in TMEM)' if a.tc else ('fp32' if a.fp32 else 'bf16 autocast, channels_last')}, "
              f"eval mode; {a.rounds} lockstep evaluator rounds of the first ply timed",
    "ms_per_round": 1e3 * tot / a.rounds, "ms_eval": t_eval / a.rounds, "ms_tree_kernels": t_tree / a.rounds,
    "evaluator": ("bk_conv3x3_bf16 (hand-written tcgen05/TMEM/TMA) for the input and the 40 trunk convolutions; 1x1 heads as one matmul; head arithmetic in PyTorch") if a.tc else "PyTorch/cuDNN (library model)",
    "roofline": {"bound": "tensor", "achieved": flops_per_leaf * evals / (t_eval * 1e-3) / 1e12, "peak": peaks.get("bf16_tflops_sustained"),
                 "unit": "TFLOP/s", "frac": flops_per_leaf * evals / (t_eval * 1e-3) / 1e12 / peaks.get("bf16_tflops_sustained", 1400.0),
                 "flops_per_leaf": flops_per_leaf, "note": "evaluator time only; the tree kernels add ms_tree_kernels per round"},
}))
