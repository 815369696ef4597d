"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes driver for oracle/liborc.so.

Only tests/, bench.py's cpu_baseline / ``--impl reference`` leg and __graft_entry__.smoke() may
import this module.  The product (blokus-engine_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")


def build(force: bool = False) -> str:
    """Compile the restatement with the committed Makefile (gcc only, a few seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "blokus_oracle.hpp", "mcts_oracle.hpp", "rng_oracle.hpp")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liborc.so"], check=True, capture_output=True)
    return _LIB_PATH


class OrcConfig(C.Structure):
    _fields_ = [
        ("sims_per_move", C.c_uint32),
        ("sample_moves", C.c_uint32),
        ("c_base", C.c_float),
        ("c_init", C.c_float),
        ("dirichlet_alpha", C.c_float),
        ("exploration_fraction", C.c_float),
        ("seed", C.c_uint64),
    ]


EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_float))

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_game_new.restype = C.c_void_p
        L.orc_game_clone.restype = C.c_void_p
        L.orc_game_clone.argtypes = [C.c_void_p]
        L.orc_game_free.argtypes = [C.c_void_p]
        for name in ("orc_game_apply",):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_game_place_piece.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        for name in ("orc_game_legal_tiles", "orc_game_scores", "orc_game_last_piece_lens"):
            getattr(L, name).argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.orc_game_num_placements.argtypes = [C.c_void_p]
        L.orc_game_board.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.orc_game_board_state.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.orc_game_current_player.argtypes = [C.c_void_p]
        L.orc_game_is_terminal.argtypes = [C.c_void_p]
        L.orc_game_is_player_active.argtypes = [C.c_void_p, C.c_int]
        L.orc_game_payoff.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.orc_game_anchors.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_game_pieces.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_game_history.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_game_digest.argtypes = [C.c_void_p]
        L.orc_game_digest.restype = C.c_uint64
        L.orc_playout.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_playout_batch.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
        L.orc_playout_batch.restype = C.c_double
        L.orc_selfplay_game.argtypes = [C.POINTER(OrcConfig), C.c_int, C.c_int, C.c_void_p, C.c_void_p] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 6
        L.orc_selfplay_batch.argtypes = [C.POINTER(OrcConfig), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_selfplay_batch.restype = C.c_double
        L.orc_test_game.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_test_game.restype = C.c_float
        L.orc_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.orc_playout_index.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_playout_index.restype = C.c_uint32
        L.orc_action_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_action_uniform.restype = C.c_float
        L.orc_dirichlet.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p]
        L.orc_det_log.argtypes = [C.c_double]
        L.orc_det_log.restype = C.c_double
        L.orc_det_exp.argtypes = [C.c_double]
        L.orc_det_exp.restype = C.c_double
        L.orc_splitmix64.argtypes = [C.c_uint64]
        L.orc_splitmix64.restype = C.c_uint64
        L.orc_ucb_factor_table.argtypes = [C.c_float, C.c_float, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _flat(shape):
    a = np.ascontiguousarray(np.array(shape, dtype=np.uint8))
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8)), a.shape[0], a.shape[1]


# ---- pieces.rs / board.rs surface --------------------------------------------------------------
def piece_points(t):
    return lib().orc_piece_points(t)


def piece_num_variants(t):
    return lib().orc_piece_num_variants(t)


def piece_variant(t, v):
    w, n = C.c_int(), C.c_int()
    offs = (C.c_int * 5)()
    k = lib().orc_piece_variant(t, v, C.byref(w), C.byref(n), offs)
    return {"width": w.value, "len": n.value, "offsets": list(offs[:k])}


def gen_variants_count(shape):
    a, p, r, c = _flat(shape)
    return lib().orc_gen_variants_count(p, r, c)


def variant_new(shape):
    a, p, r, c = _flat(shape)
    out = (C.c_uint8 * 128)()
    offs = (C.c_int * 8)()
    n, w = C.c_int(), C.c_int()
    ln = lib().orc_variant_new(p, r, c, out, offs, C.byref(n), C.byref(w))
    return {"variant": [bool(x) for x in out[:ln]], "offsets": list(offs[: n.value]), "width": w.value}


def _shape_op(fn, shape):
    a, p, r, c = _flat(shape)
    out = (C.c_uint8 * 64)()
    orr, oc = C.c_int(), C.c_int()
    fn(p, r, c, out, C.byref(orr), C.byref(oc))
    return [[bool(out[i * oc.value + j]) for j in range(oc.value)] for i in range(orr.value)]


def rotate(shape):
    return _shape_op(lib().orc_rotate, shape)


def flip(shape):
    return _shape_op(lib().orc_flip, shape)


def fresh_board_is_valid(player, shape, offset):
    a, p, r, c = _flat(shape)
    return bool(lib().orc_fresh_board_is_valid(player, p, r, c, offset))


def fresh_board_len():
    return lib().orc_fresh_board_len()


class Game:
    """Handle on one restated `blokus::game::Game` (blokus/src/game.rs:91-312)."""

    def __init__(self, handle=None):
        self._h = C.c_void_p(handle if handle is not None else lib().orc_game_new())

    def __del__(self):
        try:
            lib().orc_game_free(self._h)
        except Exception:
            pass

    def clone(self):
        return Game(lib().orc_game_clone(self._h))

    def apply(self, tile, piece_to_finish=None):
        return lib().orc_game_apply(self._h, int(tile), -1 if piece_to_finish is None else int(piece_to_finish)) == 0

    def place_piece(self, p, v, o):
        return lib().orc_game_place_piece(self._h, p, v, o)

    def legal_tiles(self):
        out = (C.c_int * 400)()
        n = lib().orc_game_legal_tiles(self._h, out)
        return list(out[:n])

    def num_placements(self):
        return lib().orc_game_num_placements(self._h)

    def board(self):
        a = np.zeros(400, dtype=np.uint8)
        lib().orc_game_board(self._h, a.ctypes.data_as(C.POINTER(C.c_uint8)))
        return a

    def board_state(self):
        a = np.zeros((5, 20, 20), dtype=np.uint8)
        lib().orc_game_board_state(self._h, a.ctypes.data_as(C.POINTER(C.c_uint8)))
        return a

    def current_player(self):
        return lib().orc_game_current_player(self._h)

    def is_terminal(self):
        return bool(lib().orc_game_is_terminal(self._h))

    def is_player_active(self, p):
        return bool(lib().orc_game_is_player_active(self._h, p))

    def scores(self):
        out = (C.c_int * 4)()
        lib().orc_game_scores(self._h, out)
        return list(out)

    def last_piece_lens(self):
        out = (C.c_int * 4)()
        lib().orc_game_last_piece_lens(self._h, out)
        return list(out)

    def payoff(self):
        out = (C.c_float * 4)()
        lib().orc_game_payoff(self._h, out)
        return list(out)

    def anchors(self, player=-1):
        out = (C.c_int * 400)()
        n = lib().orc_game_anchors(self._h, player, out)
        return list(out[:n])

    def pieces(self, player):
        out = (C.c_int * 21)()
        n = lib().orc_game_pieces(self._h, player, out)
        return list(out[:n])

    def history(self):
        pl = (C.c_int * 400)()
        tl = (C.c_int * 400)()
        n = lib().orc_game_history(self._h, pl, tl)
        return list(zip(pl[:n], tl[:n]))

    def digest(self):
        return lib().orc_game_digest(self._h)


def playout(seed, game_id, policy=0, max_plies=-1, want_hash=True):
    """policy 0: seeded uniform over ascending legal tiles; 1: smallest tile; 2: largest tile."""
    tiles = np.zeros(400, dtype=np.int16)
    players = np.zeros(400, dtype=np.int8)
    counts = np.zeros(400, dtype=np.int32)
    scores = np.zeros(4, dtype=np.int32)
    payoff = np.zeros(4, dtype=np.float32)
    h = C.c_uint64(0)
    n = lib().orc_playout(seed, game_id, policy, max_plies, tiles.ctypes.data, players.ctypes.data, counts.ctypes.data,
                          scores.ctypes.data, payoff.ctypes.data, C.addressof(h) if want_hash else None)
    return {"n_plies": n, "tiles": tiles[:n].copy(), "players": players[:n].copy(), "legal_counts": counts[:n].copy(),
            "scores": scores, "payoff": payoff, "hash": h.value}


def playout_batch(seed, first_game, n_games, n_threads=1, want_hash=True):
    hashes = np.zeros(n_games, dtype=np.uint64)
    plies = np.zeros(n_games, dtype=np.int32)
    scores = np.zeros((n_games, 4), dtype=np.int32)
    steps = C.c_int64(0)
    secs = lib().orc_playout_batch(seed, first_game, n_games, n_threads, int(want_hash), hashes.ctypes.data, plies.ctypes.data,
                                   scores.ctypes.data, C.addressof(steps))
    return {"seconds": secs, "steps": steps.value, "hashes": hashes, "plies": plies, "scores": scores}


def make_config(sims_per_move=800, sample_moves=30, c_base=19652.0, c_init=1.25, dirichlet_alpha=0.03,
                exploration_fraction=0.25, seed=0):
    return OrcConfig(sims_per_move, sample_moves, c_base, c_init, dirichlet_alpha, exploration_fraction, seed)


def selfplay_game(cfg, game_id, max_plies=-1, evaluator=None, root_cap=1 << 17):
    """evaluator(id, planes[5,20,20] u8) -> (policy[400], value[4]); None = fixed-prior stub."""
    players = np.zeros(400, dtype=np.int32)
    tiles = np.zeros(400, dtype=np.int32)
    root_off = np.zeros(401, dtype=np.int32)
    r_tile = np.zeros(root_cap, dtype=np.int32)
    r_vis = np.zeros(root_cap, dtype=np.uint32)
    r_w = np.zeros(root_cap, dtype=np.float32)
    r_p = np.zeros(root_cap, dtype=np.float32)
    payoff = np.zeros(4, dtype=np.float32)
    sims = C.c_int64(0)
    cb = None
    if evaluator is not None:
        def _cb(_user, gid, planes, policy, value):
            pl = np.ctypeslib.as_array(planes, shape=(2000,)).reshape(5, 20, 20)
            pol, val = evaluator(gid, pl)
            np.ctypeslib.as_array(policy, shape=(400,))[:] = np.asarray(pol, dtype=np.float32)
            np.ctypeslib.as_array(value, shape=(4,))[:] = np.asarray(val, dtype=np.float32)
        cb = EVAL_FN(_cb)
    n = lib().orc_selfplay_game(C.byref(cfg), game_id, max_plies, C.cast(cb, C.c_void_p) if cb else None, None,
                                players.ctypes.data, tiles.ctypes.data, root_off.ctypes.data, root_cap,
                                r_tile.ctypes.data, r_vis.ctypes.data, r_w.ctypes.data, r_p.ctypes.data,
                                payoff.ctypes.data, C.addressof(sims))
    if n < 0:
        raise RuntimeError("root_cap too small")
    roots = []
    for i in range(n):
        a, b = root_off[i], root_off[i + 1]
        roots.append({"tile": r_tile[a:b].copy(), "visits": r_vis[a:b].copy(), "value_sum": r_w[a:b].copy(), "prior": r_p[a:b].copy()})
    # history may be one longer than n when max_plies cut the game; only n plies were searched
    return {"n_plies": n, "players": players[:n].copy(), "tiles": tiles[:n].copy(), "roots": roots, "payoff": payoff, "sims": sims.value}


def _wrap_eval(evaluator):
    def _cb(_user, gid, planes, policy, value):
        pl = np.ctypeslib.as_array(planes, shape=(2000,)).reshape(5, 20, 20)
        pol, val = evaluator(gid, pl)
        np.ctypeslib.as_array(policy, shape=(400,))[:] = np.asarray(pol, dtype=np.float32)
        np.ctypeslib.as_array(value, shape=(4,))[:] = np.asarray(val, dtype=np.float32)
    return EVAL_FN(_cb)


def test_game(game_id, model, baseline, seed=0):
    """play_test_game (simulation.rs:298-332): model / baseline are evaluator(id, planes) -> (policy, value)."""
    players = np.zeros(400, dtype=np.int32)
    tiles = np.zeros(400, dtype=np.int32)
    n = C.c_int(0)
    m, b = _wrap_eval(model), _wrap_eval(baseline)
    score = lib().orc_test_game(game_id, seed, C.cast(m, C.c_void_p), C.cast(b, C.c_void_p), None, players.ctypes.data,
                                tiles.ctypes.data, C.addressof(n))
    return {"score": float(score), "players": players[: n.value].copy(), "tiles": tiles[: n.value].copy()}


def selfplay_batch(cfg, first_game, n_games, n_threads=1, max_plies=-1):
    sims = C.c_int64(0)
    secs = lib().orc_selfplay_batch(C.byref(cfg), first_game, n_games, n_threads, max_plies, C.addressof(sims))
    return {"seconds": secs, "sims": sims.value}


def philox(seed, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().orc_philox(seed, c0, c1, c2, c3, out)
    return list(out)


def dirichlet(seed, game, ply, n, alpha):
    out = np.zeros(n, dtype=np.float32)
    lib().orc_dirichlet(seed, game, ply, n, alpha, out.ctypes.data)
    return out


def ucb_factor_table(c_base, c_init, n):
    out = np.zeros(n, dtype=np.float32)
    lib().orc_ucb_factor_table(c_base, c_init, n, out.ctypes.data)
    return out
