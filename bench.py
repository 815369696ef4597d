#!/usr/bin/env python3
"""bench.py — headline benchmark of the Blokus self-play hot path on B200.

Workload (BASELINE.json configs[1]): 4096 lockstep random-playout games per GPU, legal-move generation +
apply only.  One "step" = one pass of the hot path over one batch: Game::reset for every game, then every
game played to the end on the device.  Metric: Blokus moves/s, one move = one Game::apply(tile) including
whatever advance_player / move generation it triggers.

  value     device-resident throughput: K steps timed with CUDA events on the library's stream, inputs in HBM
  e2e       the same metric through the C-ABI with HOST buffers: H2D of the batch's game ids from pinned
            memory, the kernels, D2H of plies + scores + packed histories into pinned memory, every step
            (double-buffered over two handles: the copy-out of one batch overlaps the playout of the next)
  roofline  dominant kernel k_playout against the roof that binds it: INT32 issue (bound "int32"), algorithmic
            lane-ops = sum of 120*C_rem per move generation, peak = profiles/r02_int_peak.json (probed
            on this pool with bk_probe_int_peak, clocks recorded) cross-checked live; `traffic` = measured DRAM
            bytes per launch (ncu); hbm_note keeps the HBM view with that measured traffic
  cpu_baseline  the CPU oracle (C++ restatement of the reference algorithm) on the host cores, bounded sample
  extra.env_sustained  the same step looped for >= 2 s with the clocks sampled inside the loop
  extra.mcts           BASELINE's second metric (configs[2]: 1024 games, 800 sims/move, fixed priors): device-timed
                       value, e2e (finished-game tuples read back to host memory inside the timed region, bytes
                       declared) and a CPU sample on the SAME workload (complete games of the first ids)
  extra.config5_shard  configs[4]'s per-GPU share, 8192 games to completion; at N > 1 rank 0 replays a slice of
                       rank 1's ids and compares traces and payoffs
  extra.leaf_eval      configs[3] building block: ResNet(20,256) on 1024 leaves per round, native evaluator

`--impl reference` times the CPU restatement alone (the Rust reference cannot be built here: no rustc; probed at
run time and recorded) for both metrics, so a ratio can be formed from two driver-run lines.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "blokus-engine_b200"))

GAMES_PER_GPU = 4096
ALGO_BYTES_PER_MOVE = 708  # 352 B state in + 352 B out + 4 B history if state round-trips HBM
SEED = 20261018
METRIC = "blokus_moves_per_sec_legal_gen_apply"
UNIT = "moves/s"


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs (profiling recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


def cpu_baseline(target_seconds: float = 12.0) -> dict:
    """The oracle's playout on all host cores over a bounded sample of the same workload."""
    from oracle import oracle as orc
    thr = host_threads()
    probe = orc.playout_batch(SEED, 0, 2 * thr, n_threads=thr, want_hash=False)
    rate = probe["steps"] / max(probe["seconds"], 1e-9)
    n_games = int(min(GAMES_PER_GPU, max(2 * thr, target_seconds * rate / 273.0)))
    res = orc.playout_batch(SEED, 0, n_games, n_threads=thr, want_hash=False)
    return {"value": res["steps"] / res["seconds"], "unit": UNIT, "cores": thr, "kind": "port",
            "sample": f"{n_games} of the {GAMES_PER_GPU} games (ids 0..{n_games - 1}, same seed) played to the end by the "
                      f"C++ restatement of the reference algorithm (oracle/), one game per thread on {thr} threads; "
                      f"{res['steps']} moves in {res['seconds']:.2f} s"}


MCTS_GAMES = 1024
MCTS_CFG = dict(sims_per_move=800, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.03,
                exploration_fraction=0.25, seed=SEED)


def mcts_measure(local_rank: int, rank: int, plies: int):
    """Secondary metric (BASELINE.json configs[2]): 1024 games/GPU, 800 sims/move, fixed uniform priors,
    Dirichlet alpha 0.03 / frac 0.25, the whole self-play loop on the device.  plies < 0: complete games.
    Two measurements of the same workload: device-timed (CUDA events round the one kernel) and end to end through
    the C ABI — reset, the kernel, and the finished-game training tuples (packed policy records, packed histories,
    plies, scores, payoffs) copied into HOST memory, all inside the host-timed region."""
    import numpy as np
    import torch
    from blokus_self_play import SelfPlay, Config
    sp = SelfPlay(MCTS_GAMES, Config(**MCTS_CFG), first_game_id=rank * MCTS_GAMES, device=local_rank)
    sp.run_stub(2)                      # warm-up (2 plies), then start over
    sp.reset()
    c0 = sp.counters()
    ms = sp.run_stub(plies)
    c1 = sp.counters()
    out = {k: c1[k] - c0[k] for k in c1}
    out["kernel_ms"] = ms
    out["plies"] = int(sum(len(h) for h in sp.env.history()))
    out["finished_games"] = int(sp.env.is_terminal().sum())
    # end to end: the call a user of the drop-in makes (play_training_games without the per-ply Python tuples)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sp.reset()
    sp.run_stub(plies)
    ply_off, ply_ptr, tiles, visits = sp.policy_records_packed()
    plies_h, scores_h, hist_h = sp.env.fetch()
    pay = sp.env.payoff()
    out["e2e_s"] = time.perf_counter() - t0
    out["e2e_sims"] = sp.counters()["sims"] - c1["sims"]
    out["e2e_d2h_bytes"] = int(ply_off.nbytes + ply_ptr.nbytes + tiles.nbytes + visits.nbytes + plies_h.nbytes + scores_h.nbytes
                               + hist_h.nbytes + pay.nbytes)
    out["e2e_h2d_bytes"] = 48           # the search configuration + first game id (bk_config)
    out["e2e_policy_entries"] = int(len(tiles))
    sp.close()
    return out


def config5_shard_measure(local_rank: int, rank: int, world: int, dist, games: int = 8192, check: int = 32):
    """BASELINE.json configs[4]'s per-GPU share: `games` complete self-play games (800 sims/move, stub evaluator) in one
    launch.  At N > 1 rank 0 then replays the first `check` ids of RANK 1's range as a separate small batch and compares
    action traces and payoffs with what rank 1 produced inside its shard (results depend on the global id only)."""
    import numpy as np
    import torch
    from blokus_self_play import SelfPlay, Config
    try:
        cfg = Config(**MCTS_CFG)
        first = rank * games
        sp = SelfPlay(games, cfg, first_game_id=first, device=local_rank, max_children_per_game=0)
        sp.run_stub(1)
        sp.reset()
        c0 = sp.counters()
        ms = sp.run_stub(-1)
        c1 = sp.counters()
        finished = int(sp.env.is_terminal().sum())
        hists = sp.env.history()
        pay = sp.env.payoff()
        sp.close()

        def pack(hs, py):
            m = np.zeros((check, 364), dtype=np.int32)
            for i in range(check):
                t = [p * 512 + tl for p, tl in hs[i]]
                m[i, :len(t)] = t
                m[i, 360:364] = np.round(np.asarray(py[i]) * 12).astype(np.int32)
            return m
        owner = 1 if world > 1 else 0
        mine = pack(hists[:check], pay[:check]) if rank == owner else np.zeros((check, 364), dtype=np.int32)
        if world > 1:
            t = torch.from_numpy(mine).cuda()
            dist.broadcast(t, src=owner)
            mine = t.cpu().numpy()
        same = None
        if rank == 0:
            sp2 = SelfPlay(check, cfg, first_game_id=owner * games, device=local_rank)
            sp2.run_stub(-1)
            same = bool(np.array_equal(pack(sp2.env.history(), sp2.env.payoff()), mine))
            sp2.close()
        return {"games": games, "finished": finished, "sims": c1["sims"] - c0["sims"], "kernel_ms": ms,
                "plies": int(sum(len(h) for h in hists)), "identical": same, "owner": owner, "check": check}
    except Exception as e:                      # pragma: no cover - informational block only; never fails the bench
        return {"error": str(e)[:200], "games": games, "finished": 0, "sims": 0, "kernel_ms": 0.0, "plies": 0, "identical": None}


def leaf_eval_measure(local_rank: int, rounds: int = 10, warmup: int = 3) -> dict:
    """BASELINE.json configs[3] building block: one evaluator round = ResNet(20,256) forward on 1024 leaves.
    Hand-written tcgen05 trunk (bk_conv3x3_bf16) and, beside it, the PyTorch/cuDNN bf16 path."""
    import torch
    from blokus_self_play.resnet import ResNet, LeafEvaluator
    from blokus_self_play.tc_resnet import TensorCoreLeafEvaluator
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(SEED)
    model = ResNet(20, 256).to(dev)
    planes = (torch.rand((MCTS_GAMES, 5, 20, 20), device=dev) < 0.15).float()
    out = {}
    for name, ev in (("tcgen05", TensorCoreLeafEvaluator(model)), ("cudnn_bf16", LeafEvaluator(model, bf16=True))):
        for _ in range(warmup):
            ev(planes)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(rounds):
            ev(planes)
        e1.record()
        torch.cuda.synchronize(dev)
        out[name] = e0.elapsed_time(e1) / rounds
    return out


def rust_toolchain_probe() -> dict:
    """if a Rust toolchain ever appears on the box the unmodified `blokus` crate could be the CPU arm.
    Probed at run time and recorded in the line; absent in every image seen so far."""
    import shutil
    return {"rustc": shutil.which("rustc"), "cargo": shutil.which("cargo")}


def mcts_cpu_baseline(games: int = 0) -> dict:
    """The SAME workload as the GPU's configs[2] measurement, on the host: COMPLETE games of the first `games` global
    ids of the batch (same seed, same config: 800 sims/move, stub evaluator) by the C++ restatement of
    self_play/src/simulation.rs, one game per thread (the reference's own parallelism: one process per game)."""
    from oracle import oracle as orc
    thr = host_threads()
    games = games or thr            # one complete game per host thread (16 on the single-GPU boxes, 32 on the 8-GPU ones): one wave
    cfg = orc.make_config(**{**MCTS_CFG, "c_base": 19652.0})
    r = orc.selfplay_batch(cfg, 0, games, n_threads=thr, max_plies=-1)
    return {"value": r["sims"] / r["seconds"], "unit": "sims/s", "cores": thr, "kind": "port",
            "sample": f"COMPLETE games of global ids 0..{games - 1} of the {MCTS_GAMES}-game config-3 batch (800 sims/move, stub "
                      f"evaluator, same seed) by the C++ restatement of self_play/src/simulation.rs, one game per thread on "
                      f"{thr} threads; {r['sims']} sims in {r['seconds']:.1f} s"}


def run_reference(args) -> int:
    """--impl reference: the reference's CPU algorithm (oracle port; no Rust toolchain here) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    thr = host_threads()
    games_per_step = 8 * thr
    for w in range(args.warmup):
        orc.playout_batch(SEED, 10_000_000 + w * games_per_step, max(thr, games_per_step // 4), n_threads=thr, want_hash=False)
    steps = 0
    secs = 0.0
    for k in range(args.steps):
        r = orc.playout_batch(SEED, k * games_per_step, games_per_step, n_threads=thr, want_hash=False)
        steps += r["steps"]
        secs += r["seconds"]
    value = steps / max(secs, 1e-9)
    sample = (f"each step = {games_per_step} games (of the {GAMES_PER_GPU}-game batch) played to the end by the C++ "
              f"restatement of blokus/src/*.rs on {thr} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "configs[1]: lockstep random-playout games, legal-move gen + apply only (CPU sample)",
                   "games_per_step": games_per_step, "seed": SEED},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Rust reference not buildable in this image (no rustc/cargo); this is oracle/ — a C++ restatement "
                "of the reference algorithm keeping its data structures",
        "rust_toolchain": rust_toolchain_probe(),
        "extra": {},
    }
    if not args.no_mcts:
        m = mcts_cpu_baseline(args.mcts_cpu_games)
        line["extra"]["mcts"] = {"metric": "mcts_sims_per_sec_fixed_priors", "value": m["value"], "unit": "sims/s",
                                 "config": "configs[2] (CPU sample of the same workload)", "cpu_baseline": m,
                                 "e2e": {"value": m["value"], "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU, help="games per GPU (default: configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mcts", action="store_true", help="skip the secondary MCTS sims/s measurement")
    ap.add_argument("--mcts-plies", type=int, default=-1, help="plies per game in the MCTS measurement (<0: whole games)")
    ap.add_argument("--mcts-cpu-games", type=int, default=0, help="complete games of the CPU MCTS sample (0: one per host thread)")
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the extra.env_sustained loop")
    ap.add_argument("--no-config5", action="store_true", help="skip extra.config5_shard (8192 complete games per GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    from blokus_self_play import GameBatch, probe_int_peak, PLAYOUT_NEW_GAME

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    n = args.games
    first_id = rank * n                      # games sharded by GLOBAL id; no data-path collective
    batch = GameBatch(n, device=local_rank)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    # pinned host buffers of the e2e path, double-buffered: two handles (each with its own stream), so the
    # device->host copy of one batch overlaps the playout of the next; every step still moves its ids in and
    # its results out inside the timed region
    batch2 = GameBatch(n, device=local_rank)
    lanes = []
    for b in (batch, batch2):
        ids_h = torch.arange(first_id, first_id + n, dtype=torch.int64).to(torch.int32).pin_memory()
        plies_h = torch.empty(n, dtype=torch.int32).pin_memory()
        scores_h = torch.empty((n, 4), dtype=torch.int32).pin_memory()
        hist_h = torch.empty((n, 360), dtype=torch.int16).pin_memory()
        lanes.append({"b": b, "t": (ids_h, plies_h, scores_h, hist_h),
                      "p": tuple(C.c_void_p(t.data_ptr()) for t in (ids_h, plies_h, scores_h, hist_h)), "busy": False})
    h2d_bytes = n * 4
    d2h_bytes = n * 4 + n * 4 * 4 + n * 360 * 2

    def device_step(k: int):
        # Game::reset + the playout to the end of every game, ONE launch (BK_PLAYOUT_NEW_GAME: the reset runs inside k_playout)
        batch.lib.check(batch.lib.bk_env_playout(batch._h, SEED + k, first_id, -1, PLAYOUT_NEW_GAME))

    def e2e_collect(lane) -> int:
        """wait for the lane's outstanding step and read its result from the pinned host buffers"""
        if not lane["busy"]:
            return 0
        lane["b"].sync()
        lane["busy"] = False
        return int(lane["t"][1].sum().item())

    def e2e_step(k: int) -> int:
        lane = lanes[k & 1]
        done = e2e_collect(lane)           # step k-2's results (the other lane is still in flight)
        ids_p, plies_p, scores_p, hist_p = lane["p"]
        lane["b"].run_playout_raw(SEED + k, ids_p, flags=PLAYOUT_NEW_GAME)   # H2D of the ids + reset + playout (one launch), enqueued
        lane["b"].fetch_raw_async(plies_p, scores_p, hist_p)   # D2H of plies + scores + histories, enqueued
        lane["busy"] = True
        return done

    def e2e_drain() -> int:
        return sum(e2e_collect(l) for l in lanes)

    # ---- warm-up -------------------------------------------------------------------------------
    for w in range(args.warmup):
        device_step(1000 + w)
        e2e_step(1000 + w)
    e2e_drain()

    # ---- device-resident timed region ------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    region_ms = 0.0
    kernel_ms = 0.0
    moves = 0
    lane_ops = 0
    movegens = 0
    for k in range(args.steps):
        flush_l2()
        batch.event_record(0)
        device_step(k)
        batch.event_record(1)
        region_ms += batch.event_elapsed_ms()
        kernel_ms += batch.last_kernel_ms()
        c = batch.counters()
        moves += c["total_steps"]
        lane_ops += c["lane_ops"]
        movegens += c["movegens"]
    barrier()

    # ---- the same step looped for >= 2 s of timed work, clocks sampled INSIDE the loop (sustained clocks) ----
    n_sus = torch.tensor([max(args.steps, int(args.sustained_seconds * 1e3 / max(region_ms / args.steps, 1e-3)) + 1)], device="cuda")
    if dist is not None:
        dist.all_reduce(n_sus, op=dist.ReduceOp.MAX)
    sus_steps = int(n_sus.item())
    sus_sampler = ClockSampler(local_rank)
    sus_sampler.start()
    sus_ms = 0.0
    sus_kernel_ms = 0.0
    sus_moves = 0
    sus_lane_ops = 0
    for k in range(sus_steps):
        flush_l2()
        batch.event_record(0)
        device_step(5000 + k)
        batch.event_record(1)
        sus_ms += batch.event_elapsed_ms()
        sus_kernel_ms += batch.last_kernel_ms()
        c = batch.counters()
        sus_moves += c["total_steps"]
        sus_lane_ops += c["lane_ops"]
    sus_clocks = sus_sampler.stop()
    barrier()
    burst_clocks = sampler.stop()           # "clocks": every sample taken during the value's steps and the sustained loop
    sampler = ClockSampler(local_rank)      # the remaining regions (e2e, MCTS, config 5, leaf evaluation)
    sampler.start()

    # ---- end-to-end timed region (host buffers in and out, every step) -----------------------------
    for w in range(args.warmup):
        e2e_step(2000 + w)
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    e2e_moves = 0
    per_step = []
    for k in range(args.steps):
        ts = time.perf_counter()
        e2e_moves += e2e_step(k)
        per_step.append(time.perf_counter() - ts)
    e2e_moves += e2e_drain()            # the last two steps' results land before the clock stops
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if os.environ.get("BK_DEBUG"):
        print("e2e per-step us:", [round(1e6 * x) for x in per_step], file=sys.stderr)
    barrier()

    # ---- secondary metric: MCTS sims/s (configs[2]), configs[4]'s per-GPU share, leaf evaluation ----------------
    mcts = None
    c5 = None
    leaf = None
    if not args.no_mcts:
        batch.close()
        batch2.close()
        del flush
        torch.cuda.empty_cache()
        barrier()
        mcts = mcts_measure(local_rank, rank, args.mcts_plies)
        barrier()
        if not args.no_config5:
            c5 = config5_shard_measure(local_rank, rank, world, dist)
            barrier()
        leaf = leaf_eval_measure(local_rank) if rank == 0 else None
    clocks = sampler.stop()     # sampled over the e2e steps, MCTS, config-5 shard and leaf evaluation regions

    # ---- reduce over ranks: time = max, work = sum -------------------------------------------------
    z = lambda d, k: (d[k] if d else 0)
    stats = torch.tensor([region_ms, e2e_s, kernel_ms, z(mcts, "kernel_ms"), z(mcts, "e2e_s"), z(c5, "kernel_ms"), sus_ms, sus_kernel_ms],
                         dtype=torch.float64, device="cuda")
    work = torch.tensor([moves, e2e_moves, lane_ops, movegens, z(mcts, "sims"), z(mcts, "plies"), z(mcts, "finished_games"),
                         z(mcts, "lane_ops"), z(mcts, "entries"), z(mcts, "e2e_sims"), z(mcts, "e2e_d2h_bytes"), z(c5, "games"),
                         z(c5, "finished"), z(c5, "sims"), z(c5, "plies"), sus_moves, sus_lane_ops], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    region_ms, e2e_s, kernel_ms, mcts_ms, mcts_e2e_s, c5_ms, sus_ms, sus_kernel_ms = stats.tolist()
    (moves, e2e_moves, lane_ops, movegens, mcts_sims, mcts_plies, mcts_done, mcts_lane_ops, mcts_entries, mcts_e2e_sims, mcts_d2h,
     c5_games, c5_finished, c5_sims, c5_plies, sus_moves, sus_lane_ops) = work.tolist()

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        value = moves / (region_ms * 1e-3)
        # roofline of the dominant kernel (k_playout), per launch on one GPU.  The roof that binds it is INT32 issue.
        launch_ms = kernel_ms / args.steps
        moves_per_launch = moves / args.steps / world
        lane_ops_per_launch = lane_ops / args.steps / world
        prof = {}
        for name in ("r02_k_playout_dram_bytes.json", "r02_int_peak.json"):
            try:
                prof[name] = json.load(open(os.path.join(ROOT, "profiles", name)))
            except Exception:
                prof[name] = {}
        traffic = prof["r02_k_playout_dram_bytes.json"].get("dram_bytes_per_launch")
        int_peak_live = probe_int_peak(local_rank)
        int_peak = prof["r02_int_peak.json"].get("lane_ops_per_s") or int_peak_live
        int_ach = lane_ops_per_launch / (launch_ms * 1e-3)
        sus_launch_ms = sus_kernel_ms / max(sus_steps, 1)
        sus_int_ach = (sus_lane_ops / max(sus_steps, 1) / world) / (sus_launch_ms * 1e-3)
        hbm_algo = ALGO_BYTES_PER_MOVE * moves_per_launch
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": region_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": "configs[1]: 4096 lockstep random-playout games per GPU, legal-move gen + apply only, "
                                   "every game reset and played to the end each step (the reset runs inside the playout launch)",
                       "games_per_gpu": n, "moves_per_step": moves / args.steps, "seed": SEED,
                       "l2": "flushed between timed iterations (256 MiB memset)",
                       "sharding": "global game ids, rank r owns [r*n, (r+1)*n); no data-path collective"},
            # kernels of this library launched inside the timed regions: device-resident steps (k_playout), the
            # sustained loop (the same), end-to-end steps (k_playout + k_scores); the other regions add theirs below
            "gpu_launches": args.steps + int(sus_steps) + 2 * args.steps,
            "roofline": {"bound": "int32", "achieved": int_ach, "peak": int_peak, "unit": "lane-ops/s", "frac": int_ach / int_peak,
                         "traffic": traffic, "kernel": "k_playout", "launch_ms": launch_ms,
                         "algorithmic_lane_ops_per_launch": lane_ops_per_launch,
                         "peak_source": ("profiles/r02_int_peak.json (bk_probe_int_peak on this pool's B200, LOP3/SHF mix, clocks recorded beside it)"
                                         if prof["r02_int_peak.json"] else "measured live (bk_probe_int_peak)") + "; not in MEASURED_PEAKS.json, "
                                        "nominal 148 SM x 4 x 16 lanes x 1.965 GHz = 1.86e13",
                         "peak_live": int_peak_live,
                         "sustained": {"achieved": sus_int_ach, "frac": sus_int_ach / int_peak, "launch_ms": sus_launch_ms,
                                       "see": "extra.env_sustained"},
                         "note": "algorithmic lane-ops = sum over turn-start move generations of 120*C_rem (SURVEY §8d), counted "
                                 "exactly on the device; `traffic` = measured DRAM bytes per launch (ncu --set full, profiles/)"},
            "hbm_note": {"bound": "hbm", "algorithmic_bytes_per_launch": hbm_algo, "measured_traffic_bytes_per_launch": traffic,
                         "achieved_measured_gbs": (traffic / (launch_ms * 1e-3) / 1e9) if traffic else None,
                         "notional_gbs_at_708_B_per_move": hbm_algo / (launch_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                         "peak_source": peak_kind,
                         "note": "SURVEY §8d's 708 B/move is what a design that round-trips the state through HBM every move would "
                                 "move; the persistent kernel keeps the state in registers and moves ~0.3 % of that — HBM does not bind"},
            "e2e": {"value": e2e_moves / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "clocks": burst_clocks,
            "extra": {"movegens_per_move": movegens / max(moves, 1), "kernel_only_moves_per_s": moves / (kernel_ms * 1e-3),
                      "env_sustained": {"value": sus_moves / (sus_ms * 1e-3), "unit": UNIT, "steps": int(sus_steps),
                                        "seconds_of_timed_work": sus_ms * 1e-3, "ms_per_step": sus_ms / max(sus_steps, 1),
                                        "int32_frac": sus_int_ach / int_peak, "clocks": sus_clocks,
                                        "what": "the value's step (reset + playout of the batch, L2 flushed before each) repeated until the "
                                                "CUDA-event time sums to >= 2 s; nvidia-smi sampled every 100 ms inside the loop"},
                      "clocks_other_regions": clocks},
        }
        if mcts:
            line["gpu_launches"] += 2 + 2
            line["extra"]["mcts"] = {
                "metric": "mcts_sims_per_sec_fixed_priors", "value": mcts_sims / (mcts_ms * 1e-3), "unit": "sims/s",
                "config": "configs[2]: 1024 games per GPU, 800 sims/move, fixed uniform priors (stub evaluator), Dirichlet "
                          "alpha 0.03 frac 0.25, sample_moves 30; " + ("complete games" if args.mcts_plies < 0 else f"first {args.mcts_plies} plies"),
                "sims": mcts_sims, "plies_searched": mcts_plies, "finished_games": mcts_done, "kernel_ms": mcts_ms,
                "kernel": "k_selfplay_stub_pipe at <= 9 games per SM (two warps per game), k_selfplay_stub<MINB> above — one launch: every ply's root expansion, noise, 800 sims, action, apply",
                "e2e": {"value": mcts_e2e_sims / mcts_e2e_s, "unit": "sims/s", "seconds": mcts_e2e_s,
                        "h2d_bytes_per_step": mcts["e2e_h2d_bytes"], "d2h_bytes_per_step": mcts_d2h / world,
                        "what": "bk_selfplay_reset + bk_selfplay_run_stub + the finished-game training tuples (packed policy records, "
                                "packed histories, plies, scores, payoffs) copied into host memory, host-timed, max over ranks"},
                # ~1.6 KB algorithmic HBM bytes per simulation
                "roofline": {"bound": "hbm", "achieved": 1600.0 * mcts_sims / world / (mcts_ms * 1e-3) / 1e9,
                             "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": 1600.0 * mcts_sims / world / (mcts_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None},
                "int_roofline": {"achieved": mcts_lane_ops / world / (mcts_ms * 1e-3), "peak": int_peak, "unit": "lane-ops/s",
                                 "frac": mcts_lane_ops / world / (mcts_ms * 1e-3) / int_peak},
                "child_entries_created": mcts_entries,
            }
        if c5:
            line["gpu_launches"] += 1
            line["extra"]["config5_shard"] = {
                "config": f"configs[4] per-GPU share: {int(c5['games'])} complete self-play games per GPU ({int(c5_games)} on {world} GPU"
                          f"{'s' if world > 1 else ''}, global game ids), 800 sims/move, stub evaluator, one launch per GPU",
                "games": int(c5_games), "finished_games": int(c5_finished), "sims": c5_sims, "plies": c5_plies,
                "seconds_max_over_ranks": c5_ms * 1e-3, "games_per_s": c5_games / max(c5_ms * 1e-3, 1e-9),
                "sims_per_s": c5_sims / max(c5_ms * 1e-3, 1e-9),
                "cross_check": {"identical": c5.get("identical"), "ids": [c5["owner"] * c5["games"], c5["owner"] * c5["games"] + c5["check"] - 1],
                                "owner_rank": c5["owner"], "replayed_on_rank": 0,
                                "what": "action traces and payoffs of the slice replayed on rank 0 as a separate small batch"} if "owner" in c5 else None,
                "error": c5.get("error")}
        if leaf:
            flops = 18883996800.0 * MCTS_GAMES
            line["gpu_launches"] += 43 * 10          # pack + 41 convolutions + heads per round, 10 timed rounds
            line["extra"]["leaf_eval"] = {
                "metric": "resnet20x256_leaf_evals_per_sec", "value": MCTS_GAMES / (leaf["tcgen05"] * 1e-3), "unit": "leaves/s",
                "config": "configs[3] building block: ResNet(20,256) random init, eval mode, 1024 leaves per round, one GPU; "
                          "native evaluator (bk_evaluator_forward): plane packing, 41 hand-written tcgen05/TMEM/TMA convolutions "
                          "(bf16 operands, f32 accumulate), fused head kernel — no PyTorch ops in the round",
                "ms_per_round": leaf["tcgen05"], "cudnn_bf16_ms_per_round": leaf["cudnn_bf16"],
                "speedup_vs_cudnn_bf16": leaf["cudnn_bf16"] / leaf["tcgen05"],
                "roofline": {"bound": "tensor", "achieved": flops / (leaf["tcgen05"] * 1e-3) / 1e12,
                             "peak": peaks.get("bf16_tflops_sustained", 1400.0), "unit": "TFLOP/s",
                             "frac": flops / (leaf["tcgen05"] * 1e-3) / 1e12 / peaks.get("bf16_tflops_sustained", 1400.0),
                             "traffic": None, "note": "useful FLOPs (400 positions/image); the padded 21x21 layout executes 10 % more"},
            }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
            if mcts:
                line["extra"]["mcts"]["cpu_baseline"] = mcts_cpu_baseline(args.mcts_cpu_games)
        line["rust_toolchain"] = rust_toolchain_probe()
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
