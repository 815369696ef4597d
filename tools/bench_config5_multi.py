"""BASELINE.json configs[4]: 65 536 full self-play games sharded over the GPUs of one box (8192 per GPU at 8),
800 sims/move, stub evaluator, to completion.  Launch with torchrun (one rank per GPU); no data-path
collective — NCCL carries only the barrier, the max of the timings and the cross-check slice.

Cross-check: rank 0 replays a 32-game slice of rank 1's id range on its own GPU as a
separate small batch and compares action traces + payoffs with what rank 1 produced inside its 8192-game
shard: results depend on the global game id only, not on the GPU, the shard or the batch size.
Prints one JSON line on rank 0.  --games-total overrides 65 536 (per-GPU share = total / world)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))
import numpy as np
import torch
import torch.distributed as dist
from blokus_self_play import SelfPlay, Config, MODE_SKIP_FORCED
from blokus_self_play.shard import shard_range

ap = argparse.ArgumentParser()
ap.add_argument("--games-total", type=int, default=65536)
ap.add_argument("--sims", type=int, default=800)
ap.add_argument("--skip-forced", action="store_true", help="row f3 forced-ply shortcut (training tuples unchanged)")
ap.add_argument("--check", type=int, default=32)
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

def barrier():
    if world > 1:
        dist.barrier(device_ids=[lr])
    torch.cuda.synchronize()

first, n = shard_range(a.games_total, rank, world)
cfg = Config(sims_per_move=a.sims, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03, exploration_fraction=0.25, seed=20261018)
sp = SelfPlay(n, cfg, first_game_id=first, device=lr)
if a.skip_forced:
    sp.set_mode(MODE_SKIP_FORCED, 1)
sp.run_stub(1)          # warm-up ply, then start over
sp.reset()
barrier()
c0 = sp.counters()
ms = sp.run_stub(-1)    # every game of the shard to completion; CUDA events on the launching stream
c1 = sp.counters()
barrier()
# gather the finished-game tuples on the host (north star: the only data that leaves the GPU): packed policy
# records + packed histories + payoffs, timed on the host
import time as _time
_t = _time.perf_counter()
_ply_off, _ply_ptr, _tiles, _visits = sp.policy_records_packed()
_plies_h, _scores_h, _hist_h = sp.env.fetch()
_pay = sp.env.payoff()
gather_s = _time.perf_counter() - _t
gather_bytes = _ply_ptr.nbytes + _tiles.nbytes + _visits.nbytes + _hist_h.nbytes + _plies_h.nbytes + _scores_h.nbytes + _pay.nbytes
finished = int(sp.env.is_terminal().sum())
payoff = sp.env.payoff()
hists = sp.env.history()
plies = sum(len(h) for h in hists)

# cross-check slice: the first a.check ids of rank 1's range (or of this rank's own range at world 1)
owner = 1 if world > 1 else 0
ofirst, _ = shard_range(a.games_total, owner, world)
k = a.check
def pack(hs, pay):
    m = np.zeros((k, 364), dtype=np.int32)
    for i in range(k):
        t = [p * 512 + tl for p, tl in hs[i]]
        m[i, :len(t)] = t
        m[i, 360:364] = np.round(np.asarray(pay[i]) * 12).astype(np.int32)
    return m
mine = pack(hists[:k], payoff[:k]) if rank == owner else np.zeros((k, 364), dtype=np.int32)
same = None
if world > 1:
    t = torch.from_numpy(mine).cuda()
    dist.broadcast(t, src=owner)
    mine = t.cpu().numpy()
if rank == 0:
    sp2 = SelfPlay(k, cfg, first_game_id=ofirst, device=lr)
    if a.skip_forced:
        sp2.set_mode(MODE_SKIP_FORCED, 1)
    sp2.run_stub(-1)
    replay = pack(sp2.env.history(), sp2.env.payoff())
    same = bool(np.array_equal(replay, mine))
    sp2.close()

stats = torch.tensor([ms, gather_s * 1e3, ms + gather_s * 1e3], dtype=torch.float64, device="cuda")
gb = torch.tensor([float(gather_bytes)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(gb, op=dist.ReduceOp.SUM)
work = torch.tensor([c1["sims"] - c0["sims"], c1["applies"] - c0["applies"], n, finished, plies], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    dist.all_reduce(work, op=dist.ReduceOp.SUM)
if rank == 0:
    t = stats[0].item() * 1e-3
    sims, applies, games, fin, pl = work.tolist()
    print(json.dumps({
        "config": f"configs[4]: {int(games)} full self-play games sharded over {world} B200 ({n} per GPU, global game ids), "
                  f"{a.sims} sims/move, stub evaluator, to completion" + ("; forced-ply shortcut on" if a.skip_forced else ""),
        "n_gpus": world, "games": int(games), "finished": int(fin), "plies": int(pl), "sims": int(sims),
        "seconds_max_over_ranks": t, "games_per_s": games / t, "sims_per_s": sims / t, "moves_per_s": applies / t,
        "host_gather": {"seconds_max_over_ranks": stats[1].item() * 1e-3, "bytes_all_ranks": gb.item(),
                        "what": "packed policy records (bk_selfplay_results_packed) + packed histories + scores + payoffs into host memory, per GPU",
                        "seconds_play_plus_gather": stats[2].item() * 1e-3, "games_per_s_incl_gather": games / (stats[2].item() * 1e-3)},
        "cross_check": {"ids": [ofirst, ofirst + k - 1], "owner_rank": owner, "replayed_on_rank": 0,
                        "identical_traces_and_payoffs": same}}), flush=True)
if world > 1:
    dist.barrier(device_ids=[lr])
    dist.destroy_process_group()
