"""blokus_self_play — B200-native drop-in for the reference's PyO3 module of the same name
(self_play/src/lib.rs:58-63) plus the `blokus` crate's game API, over the C ABI of
include/blokus_b200.h.  Hand-written sm_100a kernels do all the work; there is no CPU fallback.
"""
from ._lib import BkConfig, BkError, Lib, default_lib  # noqa: F401
from .game import Game, GameBatch  # noqa: F401
