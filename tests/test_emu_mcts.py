"""MCTS kernel-source logic checks on the CPU warp emulator (small trees; the GPU tests go to 800 sims)."""
import parity


def test_emu_selfplay_stub_alpha003(emu_lib, orc):
    parity.check_selfplay_stub(emu_lib, orc, 2, dict(sims_per_move=24, sample_moves=4, c_base=19652, c_init=1.25,
                                                    dirichlet_alpha=0.03, exploration_fraction=0.25, seed=77),
                               first_game_id=5, max_plies=8)


def test_emu_selfplay_stub_shipped_config(emu_lib, orc):
    """model/training.py:267-272 values (alpha 0.3), greedy action from the first ply (sample_moves 0)."""
    parity.check_selfplay_stub(emu_lib, orc, 1, dict(sims_per_move=16, sample_moves=0, c_base=19652, c_init=1.25,
                                                    dirichlet_alpha=0.3, exploration_fraction=0.25, seed=3),
                               first_game_id=0, max_plies=12)
