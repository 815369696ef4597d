"""ctypes binding of include/blokus_b200.h (the C ABI of the sm_100a library).

This is the stub a maintainer of the reference would add in place of the PyO3 module
(self_play/src/lib.rs:58-63).  There is NO CPU fallback: if the CUDA library has not been built, or
no CUDA device is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.normpath(os.path.join(_HERE, "..", "lib", "libblokus_b200.so"))

BOARD_TILES = 400
MAX_PLIES = 360

OK = 0
ERR_INVALID_ARG = -1
ERR_CUDA = -2
ERR_ILLEGAL_MOVE = -3
ERR_CAPACITY = -4
ERR_STATE = -5

PLAYOUT_HASH = 1
PLAYOUT_MIN_TILE = 2
PLAYOUT_MAX_TILE = 4
PLAYOUT_NEW_GAME = 8

MODE_SKIP_FORCED = 1
MODE_FORCE_MULTI_LEAF = 2
MODE_TREE_REUSE = 4


class BkError(RuntimeError):
    """A negative bk_status; `.code` holds it, the text is bk_last_error() (cf. Err(String))."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class BkConfig(C.Structure):
    """bk_config — self_play/src/simulation.rs:14-22 plus a seed."""

    _fields_ = [
        ("sims_per_move", C.c_uint32),
        ("sample_moves", C.c_uint32),
        ("c_base", C.c_float),
        ("c_init", C.c_float),
        ("dirichlet_alpha", C.c_float),
        ("exploration_fraction", C.c_float),
        ("seed", C.c_uint64),
    ]


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "bk_last_error": (C.c_char_p, []),
    "bk_version": (C.c_char_p, []),
    "bk_device_count": (C.c_int, []),
    "bk_piece_points": (C.c_int, [C.c_int]),
    "bk_piece_num_variants": (C.c_int, [C.c_int]),
    "bk_piece_variant": (C.c_int, [C.c_int, C.c_int, _P, _P, _P]),
    "bk_env_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_P)]),
    "bk_env_destroy": (None, [_P]),
    "bk_env_num_games": (C.c_int, [_P]),
    "bk_env_reset": (C.c_int, [_P]),
    "bk_env_clone": (C.c_int, [_P, C.POINTER(_P)]),
    "bk_env_apply": (C.c_int, [_P, _P, _P, _P]),
    "bk_env_place_piece": (C.c_int, [_P, _P, _P, _P, _P]),
    "bk_env_legal_mask": (C.c_int, [_P, _P]),
    "bk_env_legal_rows": (C.c_int, [_P, _P]),
    "bk_env_legal_tiles": (C.c_int, [_P, _P, _P]),
    "bk_env_board": (C.c_int, [_P, _P]),
    "bk_env_anchors": (C.c_int, [_P, C.c_int, _P]),
    "bk_env_current_player": (C.c_int, [_P, _P]),
    "bk_env_is_terminal": (C.c_int, [_P, _P]),
    "bk_env_is_player_active": (C.c_int, [_P, _P]),
    "bk_env_scores": (C.c_int, [_P, _P]),
    "bk_env_payoff": (C.c_int, [_P, _P]),
    "bk_env_board_state": (C.c_int, [_P, _P]),
    "bk_env_board_state_dev_f32": (C.c_int, [_P, _P]),
    "bk_env_history": (C.c_int, [_P, _P, _P, _P]),
    "bk_env_pieces": (C.c_int, [_P, _P]),
    "bk_env_last_piece_lens": (C.c_int, [_P, _P]),
    "bk_env_digest": (C.c_int, [_P, _P]),
    "bk_env_playout": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_int, C.c_uint32]),
    "bk_env_playout_ids": (C.c_int, [_P, C.c_uint64, _P, C.c_int, C.c_uint32]),
    "bk_env_fetch": (C.c_int, [_P, _P, _P, _P]),
    "bk_env_fetch_async": (C.c_int, [_P, _P, _P, _P]),
    "bk_env_sync": (C.c_int, [_P]),
    "bk_probe_int_peak": (C.c_int, [C.c_int, _P, _P]),
    "bk_env_playout_results": (C.c_int, [_P, _P, _P]),
    "bk_env_last_kernel_ms": (C.c_int, [_P, _P]),
    "bk_env_event_record": (C.c_int, [_P, C.c_int]),
    "bk_env_event_elapsed": (C.c_int, [_P, _P]),
    "bk_env_playout_counters": (C.c_int, [_P, _P]),
    "bk_conv3x3_bf16": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "bk_conv3x3_bf16_in": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "bk_evaluator_create": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(_P)]),
    "bk_evaluator_destroy": (None, [_P]),
    "bk_evaluator_max_rows": (C.c_int, [_P]),
    "bk_evaluator_forward": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P, _P]),
    "bk_selfplay_run_network": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "bk_eval_pack_planes": (C.c_int, [_P, C.c_int, _P, _P]),
    "bk_env_board_state_nhwc": (C.c_int, [_P, _P]),
    "bk_eval_heads": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P, _P, _P]),
    "bk_selfplay_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(BkConfig), C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "bk_selfplay_destroy": (None, [_P]),
    "bk_selfplay_reset": (C.c_int, [_P, C.c_uint32]),
    "bk_selfplay_run_stub": (C.c_int, [_P, C.c_int]),
    "bk_selfplay_begin_ply": (C.c_int, [_P]),
    "bk_selfplay_leaf_planes": (C.c_int, [_P, _P, _P]),
    "bk_selfplay_expand_backup": (C.c_int, [_P, _P, _P, _P]),
    "bk_selfplay_leaf_rows": (C.c_int, [_P, _P]),
    "bk_selfplay_end_ply": (C.c_int, [_P]),
    "bk_selfplay_set_stream": (C.c_int, [_P, _P]),
    "bk_selfplay_set_mode": (C.c_int, [_P, C.c_uint32, C.c_int]),
    "bk_selfplay_live_games": (C.c_int, [_P, _P]),
    "bk_selfplay_env": (_P, [_P]),
    "bk_selfplay_results": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P]),
    "bk_selfplay_results_sizes": (C.c_int, [_P, _P, _P, _P, _P]),
    "bk_selfplay_results_packed": (C.c_int, [_P, _P, _P, _P]),
    "bk_selfplay_last_root": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "bk_selfplay_training_sizes": (C.c_int, [_P, _P, _P]),
    "bk_selfplay_training_tensors": (C.c_int, [_P, _P, _P, _P]),
    "bk_selfplay_counters": (C.c_int, [_P, _P]),
    "bk_selfplay_counters_raw": (C.c_int, [_P, _P]),
    "bk_selfplay_probe_stats": (C.c_int, [_P, _P]),
    "bk_selfplay_last_kernel_ms": (C.c_int, [_P, _P]),
}


class Lib:
    """One loaded copy of the C-ABI library."""

    def __init__(self, path: str = DEFAULT_LIB):
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build the sm_100a extension first (python -c 'import __graft_entry__ as g; "
                "g.build()'). The B200 path has no CPU fallback."
            )
        self.path = path
        self.dll = C.CDLL(path)
        self.missing = []
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(self.dll, name)
            except AttributeError:
                self.missing.append(name)
                continue
            fn.restype = res
            fn.argtypes = args

    def __getattr__(self, name):
        return getattr(self.dll, name)

    def check(self, rc: int) -> int:
        if rc < 0:
            raise BkError(rc, self.dll.bk_last_error().decode("utf-8", "replace"))
        return rc

    def require_device(self) -> int:
        n = self.dll.bk_device_count()
        if n <= 0:
            raise RuntimeError("no CUDA device visible: the B200 path has no CPU fallback")
        return n


_default = None


def default_lib() -> Lib:
    global _default
    if _default is None:
        _default = Lib(os.environ.get("BK_LIB") or DEFAULT_LIB)     # BK_LIB: an alternative build of the same library (A/B probes)
    return _default
