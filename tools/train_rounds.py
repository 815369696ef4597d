"""The reference's training rounds (model/training.py:176-229 `main()`) on this engine, every stage on the device:

    per round:  self-play of `games` games with the CURRENT network as the leaf evaluator
                (SelfPlay.run_evaluator replaces the mp.Pool of play_training_game workers + handle_inference_batch)
                -> training tensors (bk_selfplay_training_tensors replaces save())
                -> DeviceReplayBuffer.extend (replaces torchrl's ReplayBuffer)
                -> `training_steps` Adam steps of train() (cross-entropy on the policy, MSE on the value,
                   model/training.py:122-142)

The optimiser side is NOT part of the hot path this repository replaces (out of scope); the script
exists to show the drop-in fit end to end and to time a round.  Defaults are the reference's TestConfig
(training.py:283-305: ResNet(2, 16), 10 sims/move, batch 64, 10 steps, 2 rounds).  Width-256 networks use the
hand-written tcgen05 evaluator.  Prints one JSON line per round.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
import torch
from blokus_self_play import SelfPlay, Config, MODE_SKIP_FORCED
from blokus_self_play.replay import DeviceReplayBuffer
from blokus_self_play.resnet import ResNet, LeafEvaluator

ap = argparse.ArgumentParser()
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--games", type=int, default=16)
ap.add_argument("--sims", type=int, default=10)
ap.add_argument("--depth", type=int, default=2)
ap.add_argument("--width", type=int, default=16)
ap.add_argument("--batch-size", type=int, default=64)
ap.add_argument("--training-steps", type=int, default=10)
ap.add_argument("--buffer-capacity", type=int, default=500000)
ap.add_argument("--max-plies", type=int, default=-1, help="cut the games short (smoke runs)")
ap.add_argument("--leaves", type=int, default=1)
ap.add_argument("--skip-forced", action="store_true")
a = ap.parse_args()

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = ResNet(a.depth, a.width).to(dev)
optimizer = torch.optim.Adam(model.parameters(), lr=0.01)          # training.py:172-175
policy_loss, value_loss = torch.nn.CrossEntropyLoss(), torch.nn.MSELoss()
buffer = DeviceReplayBuffer(a.buffer_capacity, a.batch_size, device=dev, seed=0)
cfg = Config(sims_per_move=a.sims, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.3, exploration_fraction=0.25, seed=1)
step = 0
for rnd in range(a.rounds):
    t0 = time.time()
    model.eval()
    if a.width == 256:
        from blokus_self_play.tc_resnet import TensorCoreLeafEvaluator
        ev = TensorCoreLeafEvaluator(model, max_rows=a.games * a.leaves)   # BatchNorm folded from the current weights
    else:
        ev = LeafEvaluator(model)
    sp = SelfPlay(a.games, cfg, first_game_id=rnd * a.games)        # fresh global ids every round
    sp.set_mode(MODE_SKIP_FORCED if a.skip_forced else 0, a.leaves)
    if a.width == 256:
        info = sp.run_network(ev, max_plies=a.max_plies)            # the whole round inside the library
        info["plies"] = max(len(h) for h in sp.env.history())
        ev.close()
    else:
        info = sp.run_evaluator(ev, max_plies=a.max_plies)
    torch.cuda.synchronize()
    t1 = time.time()
    n_new = buffer.extend_from(sp)
    sims = sp.counters()["sims"]
    sp.close()
    model.train()
    losses = []
    for _ in range(a.training_steps):                               # train(), training.py:122-142
        batch = buffer.sample()
        optimizer.zero_grad()
        policy, value = model(batch.get("states"))
        pl, vl = policy_loss(policy, batch.get("policies")), value_loss(value, batch.get("scores"))
        (pl + vl).backward()
        optimizer.step()
        losses.append((float(pl), float(vl)))
        step += 1
    torch.cuda.synchronize()
    t2 = time.time()
    print(json.dumps({"round": rnd, "games": a.games, "plies": info["plies"], "evaluator_rounds": info["rounds"], "sims": sims,
                      "new_samples": n_new, "buffer": len(buffer), "selfplay_s": t1 - t0, "train_s": t2 - t1,
                      "policy_loss": losses[-1][0], "value_loss": losses[-1][1], "steps": step}), flush=True)
