"""Host-side mirror of the reference's self-play client (self_play/src/lib.rs:9-32,
self_play/src/simulation.rs:267-296) over the C ABI: batches of MCTS self-play games on one B200.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import BkConfig, Lib
from .game import GameBatch, _ptr


class Config:
    """The six attributes the reference's Rust side reads by name (simulation.rs:14-22; defaults are the
    shipped model/training.py:267-272 values) plus `seed`, which the reference lacks."""

    def __init__(self, sims_per_move=50, sample_moves=30, c_base=19652, c_init=1.25, dirichlet_alpha=0.3,
                 exploration_fraction=0.25, seed=0):
        self.sims_per_move = sims_per_move
        self.sample_moves = sample_moves
        self.c_base = c_base
        self.c_init = c_init
        self.dirichlet_alpha = dirichlet_alpha
        self.exploration_fraction = exploration_fraction
        self.seed = seed


def _to_bk_config(config) -> BkConfig:
    """Reads the attributes by name, as `#[derive(FromPyObject)]` does (Python ints are fine for floats)."""
    return BkConfig(int(config.sims_per_move), int(config.sample_moves), float(config.c_base), float(config.c_init),
                    float(config.dirichlet_alpha), float(config.exploration_fraction), int(getattr(config, "seed", 0)))


class TorchBuffers:
    """Where the evaluator batches live: CUDA tensors on the handle's device, on torch's current stream.  Any object
    with the same four methods can stand in (a caller that owns device memory by other means)."""

    def __init__(self, device: int):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", device)

    def stream(self) -> int:
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def zeros(self, shape):
        return self.torch.zeros(shape, dtype=self.torch.float32, device=self.dev)

    def empty(self, shape):
        t = self.torch.empty(shape, dtype=self.torch.float32, device=self.dev)
        self.torch.cuda.synchronize(self.dev)
        return t

    def f32(self, t):
        return t.to(dtype=self.torch.float32).contiguous()

    @staticmethod
    def ptr(t) -> int:
        return t.data_ptr()

    @staticmethod
    def to_host(t) -> np.ndarray:
        return t.detach().cpu().numpy()

    def from_host(self, a: np.ndarray):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.dev)


class SelfPlay:
    """n self-play clients (bk_selfplay).  Game g has global id first_game_id + g."""

    def __init__(self, n_games: int, config, first_game_id: int = 0, device: int = 0, lib: Optional[Lib] = None,
                 max_children_per_game: int = 0):
        self.lib = lib or _lib.default_lib()
        self.lib.require_device()
        self.n = int(n_games)
        self.config = config
        self.first_game_id = first_game_id
        self._cfg = _to_bk_config(config)
        h = C.c_void_p()
        self.lib.check(self.lib.bk_selfplay_create(self.n, device, C.byref(self._cfg), first_game_id,
                                                   max_children_per_game, C.byref(h)))
        self._h = h
        self.leaves_per_round = 1
        self.mode = 0
        self.env = GameBatch(self.n, device=device, lib=self.lib, _handle=self.lib.bk_selfplay_env(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None:
            self.lib.bk_selfplay_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, first_game_id: Optional[int] = None) -> None:
        if first_game_id is not None:
            self.first_game_id = first_game_id
        self.lib.check(self.lib.bk_selfplay_reset(self._h, self.first_game_id))

    def run_stub(self, max_plies: int = -1) -> float:
        """training_game() with the fixed-prior stub evaluator, on the device; returns kernel ms."""
        self.lib.check(self.lib.bk_selfplay_run_stub(self._h, max_plies))
        return self.last_kernel_ms()

    def set_mode(self, flags: int = 0, leaves_per_round: int = 1) -> None:
        """Opt-in throughput modes (bk_selfplay_set_mode): flags = MODE_SKIP_FORCED (single-legal-tile roots are
        not searched; the training tuple is unchanged) and leaves_per_round K > 1 (K simulations per game in flight
        per evaluator round, virtual loss; visit counts then differ from the reference's).  Defaults = exact mode."""
        self.lib.check(self.lib.bk_selfplay_set_mode(self._h, int(flags), int(leaves_per_round)))
        self.mode = int(flags)
        self.leaves_per_round = int(leaves_per_round)

    # ---- external evaluator (simulation.rs:50-57 replaced by one device batch per round) ---------------
    def set_stream(self, cuda_stream: int) -> None:
        self.lib.check(self.lib.bk_selfplay_set_stream(self._h, C.c_void_p(cuda_stream)))

    def begin_ply(self) -> None:
        self.lib.check(self.lib.bk_selfplay_begin_ply(self._h))

    def leaf_planes(self, dev_planes_ptr: int, want_count: bool = True) -> int:
        cnt = C.c_int32(0)
        self.lib.check(self.lib.bk_selfplay_leaf_planes(self._h, C.c_void_p(dev_planes_ptr), C.byref(cnt) if want_count else None))
        return cnt.value

    def leaf_rows(self) -> int:
        """Rows of the evaluator batch written by the last leaf_planes(): n in the exact mode, the number of leaves
        outstanding (dense rows) in the multi-leaf mode."""
        v = C.c_int32(0)
        self.lib.check(self.lib.bk_selfplay_leaf_rows(self._h, C.byref(v)))
        return v.value

    def expand_backup(self, dev_policy_ptr: int, dev_value_ptr: int, want_count: bool = True) -> int:
        cnt = C.c_int32(0)
        self.lib.check(self.lib.bk_selfplay_expand_backup(self._h, C.c_void_p(dev_policy_ptr), C.c_void_p(dev_value_ptr),
                                                          C.byref(cnt) if want_count else None))
        return cnt.value

    def end_ply(self) -> None:
        self.lib.check(self.lib.bk_selfplay_end_ply(self._h))

    def run_evaluator(self, evaluator: Callable, max_plies: int = -1, buffers=None) -> dict:
        """training_game() for every client with a caller-supplied evaluator.

        evaluator(planes[R,5,20,20] float32) -> (policy[R,400] float32 in the mover's frame, value[R,4] float32
        in relative-seat order), exactly the contract of the reference's inference server
        (model/training.py:43-67, model/resnet.py:69-94), but on ONE contiguous device batch.  Rows are DENSE in
        every mode: only positions that wait for an answer are written and evaluated (R = the live games with a
        leaf out in the exact mode, the leaves outstanding in the multi-leaf mode), in (game, slot) order.
        `buffers` owns the batch memory (default: CUDA tensors on this handle's device, TorchBuffers)."""
        n = self.n * self.leaves_per_round      # capacity of the evaluator batch
        buf = buffers or TorchBuffers(self.env.device)
        planes = buf.zeros((n, 5, 20, 20))
        if hasattr(buf, "stream"):
            self.set_stream(buf.stream())
        rounds = 0
        plies = 0
        while (max_plies < 0 or plies < max_plies) and self.live_games() > 0:
            self.begin_ply()
            pending = self.leaf_planes(buf.ptr(planes))
            while pending > 0:
                rows = self.leaf_rows()
                if rows > 0:
                    policy, value = evaluator(planes[:rows])
                    policy, value = buf.f32(policy), buf.f32(value)
                else:                                   # only kept trees resuming: nothing to evaluate this round
                    policy, value = planes[:1, 0].reshape(-1)[:400], planes[:1, 0].reshape(-1)[:4]
                    policy, value = buf.f32(policy), buf.f32(value)
                pending = self.expand_backup(buf.ptr(policy), buf.ptr(value))
                rounds += 1
                if pending > 0:
                    self.leaf_planes(buf.ptr(planes), want_count=False)
            self.end_ply()
            plies += 1
        return {"plies": plies, "rounds": rounds}

    def run_network(self, evaluator, max_plies: int = -1) -> dict:
        """training_game() for every client with a native network evaluator (tc_resnet.TensorCoreLeafEvaluator, a
        `bk_evaluator`): the whole round — planes, network, expand + backup — stays inside the library
        (bk_selfplay_run_network); the host reads 8 bytes per round."""
        evaluator.reserve(self.n * self.leaves_per_round)
        rounds, evals = C.c_int64(0), C.c_int64(0)
        self.lib.check(self.lib.bk_selfplay_run_network(self._h, evaluator.handle, int(max_plies), C.byref(rounds), C.byref(evals)))
        return {"rounds": rounds.value, "evals": evals.value, "ms": self.last_kernel_ms()}

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0)
        self.lib.check(self.lib.bk_selfplay_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def live_games(self) -> int:
        v = C.c_int32(0)
        self.lib.check(self.lib.bk_selfplay_live_games(self._h, C.byref(v)))
        return v.value

    def counters(self) -> dict:
        c = np.zeros(6, dtype=np.uint64)
        self.lib.check(self.lib.bk_selfplay_counters(self._h, _ptr(c)))
        return {"sims": int(c[0]), "applies": int(c[1]), "movegens": int(c[2]), "lane_ops": int(c[3]),
                "entries": int(c[4]), "nodes": int(c[5])}

    def policy_records(self):
        """Per game, per searched ply: (tiles int16[k], visits uint32[k]) of the root's children (views into one packed
        gather, bk_selfplay_results_packed)."""
        ply_off, ply_ptr, tiles, visits = self.policy_records_packed()
        out = []
        for g in range(self.n):
            a, b = int(ply_off[g]), int(ply_off[g + 1])
            out.append([(tiles[ply_ptr[k]:ply_ptr[k + 1]], visits[ply_ptr[k]:ply_ptr[k + 1]]) for k in range(a, b)])
        return out

    def policy_records_unpacked(self, policy_cap: int = 32768):
        """The same records through bk_selfplay_results (two copies per game into [n][policy_cap] host arrays)."""
        plies = np.zeros(self.n, dtype=np.int32)
        off = np.zeros((self.n, _lib.MAX_PLIES + 1), dtype=np.int32)
        tiles = np.zeros((self.n, policy_cap), dtype=np.int16)
        visits = np.zeros((self.n, policy_cap), dtype=np.uint32)
        self.lib.check(self.lib.bk_selfplay_results(self._h, _ptr(plies), _ptr(off), policy_cap, _ptr(tiles), _ptr(visits)))
        out = []
        for g in range(self.n):
            recs = []
            for k in range(int(plies[g])):
                a, b = int(off[g, k]), int(off[g, k + 1])
                recs.append((tiles[g, a:b].copy(), visits[g, a:b].copy()))
            out.append(recs)
        return out

    def policy_records_packed(self):
        """All policy records as one compressed-row structure (bk_selfplay_results_packed): (ply_offset int64[n+1],
        ply_ptr int64[P+1], tiles int16[E], visits uint32[E]); ply k of game g owns entries
        ply_ptr[ply_offset[g]+k] : ply_ptr[ply_offset[g]+k+1].  One kernel and three copies whatever the batch size."""
        tp, te = C.c_int64(0), C.c_int64(0)
        ply_off = np.zeros(self.n + 1, dtype=np.int64)
        self.lib.check(self.lib.bk_selfplay_results_sizes(self._h, C.byref(tp), C.byref(te), _ptr(ply_off), None))
        ply_ptr = np.zeros(tp.value + 1, dtype=np.int64)
        tiles = np.zeros(max(te.value, 1), dtype=np.int16)
        visits = np.zeros(max(te.value, 1), dtype=np.uint32)
        self.lib.check(self.lib.bk_selfplay_results_packed(self._h, _ptr(ply_ptr), _ptr(tiles), _ptr(visits)))
        return ply_off, ply_ptr, tiles[: te.value], visits[: te.value]

    def last_root(self):
        """Root children of the last searched ply: per game dict(tile, visits, value_sum, prior)."""
        cnt = np.zeros(self.n, dtype=np.int32)
        tile = np.zeros((self.n, 400), dtype=np.int16)
        vis = np.zeros((self.n, 400), dtype=np.uint32)
        w = np.zeros((self.n, 400), dtype=np.float32)
        p = np.zeros((self.n, 400), dtype=np.float32)
        self.lib.check(self.lib.bk_selfplay_last_root(self._h, _ptr(cnt), _ptr(tile), _ptr(vis), _ptr(w), _ptr(p)))
        return [{"tile": tile[g, : cnt[g]].copy(), "visits": vis[g, : cnt[g]].copy(), "value_sum": w[g, : cnt[g]].copy(),
                 "prior": p[g, : cnt[g]].copy()} for g in range(self.n)]

    def training_tensors(self, buffers=None):
        """`save()` of model/training.py:70-119 on the device: (states[P,5,20,20], policies[P,400], values[P,4],
        ply_offsets[n+1]) over all searched plies P, game-major, in `buffers`' memory (default: CUDA tensors)."""
        total = C.c_int64(0)
        offs = np.zeros(self.n + 1, dtype=np.int64)
        self.lib.check(self.lib.bk_selfplay_training_sizes(self._h, C.byref(total), _ptr(offs)))
        P = max(int(total.value), 1)
        buf = buffers or TorchBuffers(self.env.device)
        st, po, va = buf.empty((P, 5, 20, 20)), buf.empty((P, 400)), buf.empty((P, 4))
        self.lib.check(self.lib.bk_selfplay_training_tensors(self._h, *[C.c_void_p(buf.ptr(t)) for t in (st, po, va)]))
        n = int(total.value)
        return st[:n], po[:n], va[:n], offs

    def game_data(self) -> List[Tuple[list, list, list]]:
        """What training_game() returns for each game (simulation.rs:293-295):
        (history [(player, tile)], policies [[(tile, prob)]], values [4]).
        The probabilities of every ply of every game are computed in one vectorised pass over the packed gather
        (f32 visits / f32 total, simulation.rs:222); only the final nesting into the reference's list-of-tuples is Python."""
        hist = self.env.history()
        pay = self.env.payoff()
        ply_off, ply_ptr, tiles, visits = self.policy_records_packed()
        n_plies = len(ply_ptr) - 1
        if n_plies > 0 and len(visits) > 0:
            starts = ply_ptr[:-1]
            totals = np.add.reduceat(visits.astype(np.uint64), np.minimum(starts, len(visits) - 1)).astype(np.float32)
            counts = np.diff(ply_ptr)
            totals[counts == 0] = 1.0
            probs = visits.astype(np.float32) / np.repeat(totals, counts)
        else:
            probs = np.zeros(0, dtype=np.float32)
        pairs = list(zip(tiles.astype(int).tolist(), probs.tolist()))
        ptr = ply_ptr.tolist()
        out = []
        for g in range(self.n):
            a, b = int(ply_off[g]), int(ply_off[g + 1])
            pols = [pairs[ptr[k]:ptr[k + 1]] for k in range(a, b)]
            if len(hist[g]) != len(pols):       # simulation.rs:293-295: one policy per history entry, by construction
                raise ValueError(f"game {g}: {len(hist[g])} plies of history but {len(pols)} searched plies — the game was "
                                 "advanced outside the search (env.apply / playout), the tuple would be misaligned")
            out.append((hist[g], pols, pay[g].tolist()))
        return out


def play_training_games(ids: Sequence[int], config, device: int = 0, lib: Optional[Lib] = None):
    """Batched form of play_training_game for the fixed-prior stub evaluator: ids must be consecutive
    global game ids.  Returns [(history, policies, values)] in id order."""
    ids = list(ids)
    if not ids:
        return []
    if ids != list(range(ids[0], ids[0] + len(ids))):
        raise ValueError("ids must be consecutive")
    sp = SelfPlay(len(ids), config, first_game_id=ids[0], device=device, lib=lib)
    try:
        sp.run_stub(-1)
        return sp.game_data()
    finally:
        sp.close()


def host_evaluator(fn: Callable):
    """Adapt fn(planes: np.ndarray[n,5,20,20]) -> (policy np[n,400], value np[n,4]) to CUDA tensors."""
    def wrapped(planes):
        import torch
        pol, val = fn(planes.detach().cpu().numpy())
        return (torch.from_numpy(np.ascontiguousarray(pol, dtype=np.float32)).to(planes.device),
                torch.from_numpy(np.ascontiguousarray(val, dtype=np.float32)).to(planes.device))
    return wrapped


def play_training_game(id: int, config, inference_queue, pipe, device: int = 0, lib: Optional[Lib] = None, buffers=None):
    """Drop-in for the reference's `play_training_game(id, config, inference_queue, pipe)`
    (self_play/src/lib.rs:9-32): one game, every leaf sent to the Python inference server with the
    reference's own protocol — `inference_queue.put((id, planes))` with planes as nested bool lists
    [5][20][20] and `pipe.recv()` -> (policy[400], value[4]) (simulation.rs:50-57).  Returns
    (history, policies, values).  This keeps `model/training.py` running unmodified; the batched
    `SelfPlay.run_evaluator` is the fast path."""
    sp = SelfPlay(1, config, first_game_id=int(id), device=device, lib=lib)
    buf = buffers or TorchBuffers(device)

    def ev(planes):
        inference_queue.put((id, buf.to_host(planes)[0].astype(bool).tolist()))
        policy, value = pipe.recv()
        return (buf.from_host(np.asarray(policy, dtype=np.float32)[None, :]),
                buf.from_host(np.asarray(value, dtype=np.float32)[None, :]))

    try:
        sp.run_evaluator(ev, -1, buffers=buf)
        return sp.game_data()[0]
    finally:
        sp.close()
