"""Host-side mirror of the `blokus` crate's game-state API (blokus/src/game.rs:91-312) over the C ABI.

`GameBatch` is n games stepped in lockstep on one B200; `Game` is the n == 1 view with the reference's
method names (reset / apply / place_piece / get_legal_tiles / get_board / get_score / get_payoff /
is_terminal / is_player_active / get_board_state / current_player / history), so a test written against
the reference's `Game` reads the same here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import BkError, Lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class GameBatch:
    """n independent Game values resident in HBM (bk_env)."""

    def __init__(self, n_games: int, device: int = 0, lib: Optional[Lib] = None, _handle=None):
        self.lib = lib or _lib.default_lib()
        self.n = int(n_games)
        self.device = device
        self._owned = _handle is None
        if _handle is None:
            self.lib.require_device()
            h = C.c_void_p()
            self.lib.check(self.lib.bk_env_create(self.n, device, C.byref(h)))
            self._h = h
        else:
            self._h = C.c_void_p(_handle)

    def close(self):
        if getattr(self, "_h", None) is not None and self._owned:
            self.lib.bk_env_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Game::reset / Clone ------------------------------------------------------------------
    def reset(self):
        self.lib.check(self.lib.bk_env_reset(self._h))

    def clone(self) -> "GameBatch":
        h = C.c_void_p()
        self.lib.check(self.lib.bk_env_clone(self._h, C.byref(h)))
        out = GameBatch.__new__(GameBatch)
        out.lib, out.n, out.device, out._owned, out._h = self.lib, self.n, self.device, True, h
        return out

    # ---- Game::apply / place_piece -------------------------------------------------------------
    def apply(self, tiles: Sequence[int], piece_to_finish: Optional[Sequence[int]] = None, strict: bool = True) -> np.ndarray:
        """Game::apply for every game (tile < 0 skips a game). Returns per-game status."""
        t = np.ascontiguousarray(np.asarray(tiles, dtype=np.int32).reshape(self.n))
        f = None if piece_to_finish is None else np.ascontiguousarray(np.asarray(piece_to_finish, dtype=np.int32).reshape(self.n))
        st = np.zeros(self.n, dtype=np.int32)
        rc = self.lib.bk_env_apply(self._h, _ptr(t), None if f is None else _ptr(f), _ptr(st))
        if rc < 0 and (strict or rc != _lib.ERR_ILLEGAL_MOVE):
            self.lib.check(rc)
        return st

    def place_piece(self, p: Sequence[int], v: Sequence[int], o: Sequence[int], strict: bool = True) -> np.ndarray:
        pa, va, oa = (np.ascontiguousarray(np.asarray(x, dtype=np.int32).reshape(self.n)) for x in (p, v, o))
        st = np.zeros(self.n, dtype=np.int32)
        rc = self.lib.bk_env_place_piece(self._h, _ptr(pa), _ptr(va), _ptr(oa), _ptr(st))
        if rc < 0 and (strict or rc != _lib.ERR_ILLEGAL_MOVE):
            self.lib.check(rc)
        return st

    # ---- accessors -------------------------------------------------------------------------------
    def _get(self, fn, shape, dtype, *extra):
        out = np.zeros(shape, dtype=dtype)
        self.lib.check(fn(self._h, *extra, _ptr(out)))
        return out

    def legal_mask(self) -> np.ndarray:
        return self._get(self.lib.bk_env_legal_mask, (self.n, 400), np.uint8)

    def legal_rows(self) -> np.ndarray:
        return self._get(self.lib.bk_env_legal_rows, (self.n, 20), np.uint32)

    def legal_tiles(self):
        """Game::get_legal_tiles per game (game.rs:242-244), ascending: the list form of the C ABI."""
        counts = np.zeros(self.n, dtype=np.int32)
        tiles = np.zeros((self.n, 400), dtype=np.int16)
        self.lib.check(self.lib.bk_env_legal_tiles(self._h, _ptr(counts), _ptr(tiles)))
        return [tiles[g, : counts[g]].astype(int).tolist() for g in range(self.n)]

    def board(self) -> np.ndarray:
        return self._get(self.lib.bk_env_board, (self.n, 400), np.uint8)

    def anchors(self, player: int = -1) -> np.ndarray:
        return self._get(self.lib.bk_env_anchors, (self.n, 400), np.uint8, player)

    def current_player(self) -> np.ndarray:
        return self._get(self.lib.bk_env_current_player, (self.n,), np.int32)

    def is_terminal(self) -> np.ndarray:
        return self._get(self.lib.bk_env_is_terminal, (self.n,), np.int32).astype(bool)

    def is_player_active(self) -> np.ndarray:
        return self._get(self.lib.bk_env_is_player_active, (self.n, 4), np.int32).astype(bool)

    def scores(self) -> np.ndarray:
        return self._get(self.lib.bk_env_scores, (self.n, 4), np.int32)

    def payoff(self) -> np.ndarray:
        return self._get(self.lib.bk_env_payoff, (self.n, 4), np.float32)

    def board_state(self) -> np.ndarray:
        return self._get(self.lib.bk_env_board_state, (self.n, 5, 20, 20), np.uint8)

    def pieces(self) -> np.ndarray:
        return self._get(self.lib.bk_env_pieces, (self.n, 4), np.uint32)

    def last_piece_lens(self) -> np.ndarray:
        return self._get(self.lib.bk_env_last_piece_lens, (self.n, 4), np.int32)

    def digest(self) -> np.ndarray:
        return self._get(self.lib.bk_env_digest, (self.n,), np.uint64)

    def history(self):
        """Game::history per game: list of (player, tile)."""
        cnt = np.zeros(self.n, dtype=np.int32)
        pl = np.zeros((self.n, _lib.MAX_PLIES), dtype=np.int32)
        tl = np.zeros((self.n, _lib.MAX_PLIES), dtype=np.int32)
        self.lib.check(self.lib.bk_env_history(self._h, _ptr(cnt), _ptr(pl), _ptr(tl)))
        return [list(zip(pl[g, : cnt[g]].tolist(), tl[g, : cnt[g]].tolist())) for g in range(self.n)]

    # ---- lockstep random playouts (BASELINE.json configs 1-2) -------------------------------------
    def playout(self, seed: int = 0, first_game_id: int = 0, max_plies: int = -1, flags: int = 0,
                game_ids: Optional[np.ndarray] = None) -> dict:
        """Play every game forward on the device (legal-move gen + seeded choice + apply per ply).
        game_ids (uint32[n], host) gives each game its global id; default first_game_id + g."""
        if game_ids is not None:
            ids = np.ascontiguousarray(np.asarray(game_ids, dtype=np.uint32).reshape(self.n))
            self.lib.check(self.lib.bk_env_playout_ids(self._h, seed, _ptr(ids), max_plies, flags))
        else:
            self.lib.check(self.lib.bk_env_playout(self._h, seed, first_game_id, max_plies, flags))
        steps = np.zeros(self.n, dtype=np.int32)
        hashes = np.zeros(self.n, dtype=np.uint64)
        self.lib.check(self.lib.bk_env_playout_results(self._h, _ptr(steps), _ptr(hashes)))
        ms = C.c_float(0)
        self.lib.check(self.lib.bk_env_last_kernel_ms(self._h, C.byref(ms)))
        ctr = np.zeros(3, dtype=np.uint64)
        self.lib.check(self.lib.bk_env_playout_counters(self._h, _ptr(ctr)))
        return {"steps": steps, "hash": hashes, "kernel_ms": ms.value, "total_steps": int(ctr[0]),
                "movegens": int(ctr[1]), "lane_ops": int(ctr[2])}


    def run_playout_raw(self, seed: int, ids_ptr, max_plies: int = -1, flags: int = 0) -> None:
        """bk_env_playout_ids with a caller-held (pinned) uint32 id buffer; no result copies."""
        self.lib.check(self.lib.bk_env_playout_ids(self._h, seed, ids_ptr, max_plies, flags))

    def fetch_raw(self, plies_ptr, scores_ptr, hist_ptr) -> None:
        """bk_env_fetch straight into caller-held (pinned) buffers."""
        self.lib.check(self.lib.bk_env_fetch(self._h, plies_ptr, scores_ptr, hist_ptr))

    def fetch_raw_async(self, plies_ptr, scores_ptr, hist_ptr) -> None:
        """bk_env_fetch_async: enqueue the gather into caller-held pinned buffers; valid after sync()."""
        self.lib.check(self.lib.bk_env_fetch_async(self._h, plies_ptr, scores_ptr, hist_ptr))

    def sync(self) -> None:
        self.lib.check(self.lib.bk_env_sync(self._h))

    def fetch(self):
        """Finished-batch gather: plies[n], scores[n,4], packed history uint16[n,360] (tile | player<<9)."""
        plies = np.zeros(self.n, dtype=np.int32)
        scores = np.zeros((self.n, 4), dtype=np.int32)
        hist = np.zeros((self.n, _lib.MAX_PLIES), dtype=np.uint16)
        self.fetch_raw(_ptr(plies), _ptr(scores), _ptr(hist))
        return plies, scores, hist

    def event_record(self, which: int) -> None:
        self.lib.check(self.lib.bk_env_event_record(self._h, which))

    def event_elapsed_ms(self) -> float:
        ms = C.c_float(0)
        self.lib.check(self.lib.bk_env_event_elapsed(self._h, C.byref(ms)))
        return ms.value

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0)
        self.lib.check(self.lib.bk_env_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def counters(self):
        ctr = np.zeros(3, dtype=np.uint64)
        self.lib.check(self.lib.bk_env_playout_counters(self._h, _ptr(ctr)))
        return {"total_steps": int(ctr[0]), "movegens": int(ctr[1]), "lane_ops": int(ctr[2])}


def probe_int_peak(device: int = 0, lib: Optional[Lib] = None) -> float:
    """Measured integer-pipe peak of the device, 32-bit lane-ops per second."""
    lib = lib or _lib.default_lib()
    v = C.c_double(0)
    ms = C.c_float(0)
    lib.check(lib.bk_probe_int_peak(device, C.byref(v), C.byref(ms)))
    return v.value


class Game:
    """One game with the reference's `Game` method names (blokus/src/game.rs)."""

    def __init__(self, device: int = 0, lib: Optional[Lib] = None, _batch: Optional[GameBatch] = None):
        self._b = _batch if _batch is not None else GameBatch(1, device=device, lib=lib)

    @staticmethod
    def reset(device: int = 0, lib: Optional[Lib] = None) -> "Game":  # game.rs:102
        return Game(device=device, lib=lib)

    def clone(self) -> "Game":
        return Game(_batch=self._b.clone())

    def apply(self, tile: int, piece_to_finish: Optional[int] = None) -> None:  # game.rs:150
        """Raises BkError("Invalid move ...") like the reference's Err — but leaves the game untouched."""
        self._b.apply([tile], None if piece_to_finish is None else [piece_to_finish])

    def place_piece(self, p: int, v: int, o: int) -> "Game":  # game.rs:116 (returns a NEW game)
        ns = self.clone()
        ns._b.place_piece([p], [v], [o])
        return ns

    def get_board(self) -> np.ndarray:  # game.rs:196
        return self._b.board()[0]

    def current_player(self) -> int:  # game.rs:225
        return int(self._b.current_player()[0])

    def get_current_anchors(self):  # game.rs:238
        return set(np.flatnonzero(self._b.anchors(-1)[0]).tolist())

    def get_legal_tiles(self):  # game.rs:242 (ascending here; arbitrary order in the reference)
        return self._b.legal_tiles()[0]

    def get_score(self):  # game.rs:247
        return self._b.scores()[0].tolist()

    def get_payoff(self):  # game.rs:252
        return self._b.payoff()[0].tolist()

    def is_terminal(self) -> bool:  # game.rs:275
        return bool(self._b.is_terminal()[0])

    def is_player_active(self, player: int) -> bool:  # game.rs:279
        return bool(self._b.is_player_active()[0, player])

    def get_board_state(self) -> np.ndarray:  # game.rs:283
        return self._b.board_state()[0].astype(bool)

    def get_current_player_pieces(self):  # game.rs:230 — ids of the remaining pieces, list order
        mask = int(self._b.pieces()[0, self.current_player()])
        return [i for i in range(21) if (mask >> i) & 1]

    def get_piece(self, player: int, piece: int, variant: int) -> dict:  # game.rs:234
        """PieceVariant of the `piece`-th REMAINING piece of `player` (board.rs:147-153 list positions):
        {"offsets": [...], "width": w, "len": variant.len(), "piece_id": id} (pieces.rs:58-98)."""
        mask = int(self._b.pieces()[0, player])
        ids = [i for i in range(21) if (mask >> i) & 1]
        if not 0 <= piece < len(ids):
            raise IndexError("piece index out of range of the player's remaining pieces")
        w, n = C.c_int(0), C.c_int(0)
        offs = (C.c_int * 5)()
        k = self._b.lib.bk_piece_variant(ids[piece], variant, C.byref(w), C.byref(n), offs)
        if k < 0:
            raise IndexError("variant index out of range")
        return {"offsets": list(offs[:k]), "width": w.value, "len": n.value, "piece_id": ids[piece]}

    @property
    def history(self):  # game.rs:94
        return self._b.history()[0]


__all__ = ["GameBatch", "Game", "BkError", "probe_int_peak"]
