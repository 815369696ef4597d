"""Arena games: the reference's `play_test_game` (self_play/src/lib.rs:34-56,
simulation.rs:233-265,298-332) on the B200 game engine.  No tree search: one evaluator call per ply; seat 0
plays the child with the highest prior (strict `>`, so the first maximum in ascending tile order), every other
seat plays a uniformly random child.  thread_rng is replaced by the seeded Philox spec (purpose 3)."""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np

from ._lib import Lib
from .game import GameBatch

_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def _philox0(seed: int, c0: int, c1: int, c2: int, c3: int) -> int:
    """First word of Philox4x32-10 (host copy of the spec in csrc/bk_rng.cuh, a few draws per ply)."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    x0, x1, x2, x3 = c0, c1, c2, c3
    for _ in range(10):
        p0, p1 = _M0 * x0, _M1 * x2
        x0, x1, x2, x3 = ((p1 >> 32) ^ x1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ x3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return x0


def _frame_index(tile: int, cur: int) -> int:
    """Index of absolute tile inside the mover-frame policy vector (inverse of simulation.rs:25-34)."""
    r, c = divmod(tile, 20)
    if cur == 0:
        return r * 20 + c
    if cur == 1:
        return (19 - c) * 20 + r
    if cur == 2:
        return (19 - r) * 20 + (19 - c)
    return c * 20 + (19 - r)


def _children(policy_frame: np.ndarray, legal: Sequence[int], cur: int):
    """evaluate()'s expansion (simulation.rs:66-81): legal tiles with p > 0, prior = exp(p) / sequential f32 sum."""
    tiles, e = [], []
    for t in legal:
        p = np.float32(policy_frame[_frame_index(t, cur)])
        if p > 0:
            tiles.append(t)
            e.append(np.float32(np.exp(np.float64(p))))
    total = np.float32(0.0)
    for x in e:
        total = np.float32(total + x)
    return tiles, [np.float32(x / total) for x in e]


def play_test_games(ids: Sequence[int], model: Callable, baseline: Callable, seed: int = 0, device: int = 0,
                    lib: Optional[Lib] = None):
    """Batched arena: model / baseline are evaluator(planes[n,5,20,20] uint8) -> (policy[n,400], value[n,4]) on host
    arrays, called once per ply on the whole batch.  Returns (scores payoff[0] per game, histories)."""
    ids = list(ids)
    n = len(ids)
    batch = GameBatch(n, device=device, lib=lib)
    try:
        while True:
            term = batch.is_terminal()
            if term.all():
                break
            planes = batch.board_state()
            cur = batch.current_player()
            legal = batch.legal_tiles()
            pol_m, _ = model(planes)
            pol_b, _ = baseline(planes)
            plies = [len(h) for h in batch.history()]
            tiles = []
            for g in range(n):
                if term[g]:
                    tiles.append(-1)
                    continue
                c = int(cur[g])
                kids, priors = _children(np.asarray(pol_m if c == 0 else pol_b)[g], legal[g], c)
                if c != 0:                                               # simulation.rs:248-253
                    idx = (_philox0(seed, ids[g] & 0xFFFFFFFF, plies[g], 3, 0) * len(kids)) >> 32
                    tiles.append(kids[idx])
                else:                                                    # simulation.rs:256-264
                    best, hi = 0, np.float32(0.0)
                    for t, p in zip(kids, priors):
                        if p > hi:
                            hi, best = p, t
                    tiles.append(best)
            batch.apply(tiles)
        return batch.payoff()[:, 0].tolist(), batch.history()
    finally:
        batch.close()


def play_test_game(id: int, model_queue, baseline_queue, pipe, seed: int = 0, device: int = 0, lib: Optional[Lib] = None) -> float:
    """Drop-in for the reference's `play_test_game(id, model_queue, baseline_queue, pipe)` -> payoff of seat 0, with
    the reference's IPC protocol: `queue.put((id, planes))` (nested bool lists) / `pipe.recv()` -> (policy, value)."""
    def via(queue):
        def ev(planes):
            queue.put((id, planes[0].astype(bool).tolist()))
            policy, value = pipe.recv()
            return np.asarray(policy, dtype=np.float32)[None, :], np.asarray(value, dtype=np.float32)[None, :]
        return ev

    # the reference sends ONE request per ply, to the queue of the seat to move (simulation.rs:311-318)
    state = {"m": via(model_queue), "b": via(baseline_queue)}

    batch = GameBatch(1, device=device, lib=lib)
    try:
        ply = 0
        while not batch.is_terminal()[0]:
            planes = batch.board_state()
            c = int(batch.current_player()[0])
            pol, _ = state["m" if c == 0 else "b"](planes)
            kids, priors = _children(pol[0], batch.legal_tiles()[0], c)
            if c != 0:
                tile = kids[(_philox0(seed, id & 0xFFFFFFFF, ply, 3, 0) * len(kids)) >> 32]
            else:
                tile, hi = 0, np.float32(0.0)
                for t, p in zip(kids, priors):
                    if p > hi:
                        hi, tile = p, t
            batch.apply([tile])
            ply += 1
        return float(batch.payoff()[0, 0])
    finally:
        batch.close()
