"""blokus_self_play — B200-native drop-in for the reference's PyO3 module of the same name
(self_play/src/lib.rs:58-63) plus the `blokus` crate's game API, over the C ABI of
include/blokus_b200.h.  Hand-written sm_100a kernels do all the work; there is no CPU fallback.
"""
from ._lib import (BkConfig, BkError, Lib, default_lib, DEFAULT_LIB, MAX_PLIES,  # noqa: F401
                   PLAYOUT_HASH, PLAYOUT_MIN_TILE, PLAYOUT_MAX_TILE, PLAYOUT_NEW_GAME, MODE_SKIP_FORCED, MODE_FORCE_MULTI_LEAF, MODE_TREE_REUSE,
                   ERR_ILLEGAL_MOVE, ERR_CUDA, ERR_INVALID_ARG, ERR_CAPACITY, ERR_STATE)
from .game import Game, GameBatch, probe_int_peak  # noqa: F401
from .selfplay import Config, SelfPlay, play_training_games, play_training_game, host_evaluator  # noqa: F401
from .arena import play_test_game, play_test_games  # noqa: F401
