"""The seeded generator SPEC that stands in for rand::thread_rng / rand_distr::Dirichlet
(self_play/src/simulation.rs:107-109,120) — checked on the oracle side; the GPU tests prove the
kernels' separate implementation agrees bit for bit (priors after noise, sampled actions)."""
import math

import numpy as np


def test_philox_known_answers(orc):
    # Random123 kat_vectors: philox4x32 10 rounds
    assert orc.philox(0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox(0xffffffffffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox((0x299f31d0 << 32) | 0xa4093822, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_det_log_exp_accuracy(orc):
    L = orc.lib()
    rng = np.random.default_rng(0)
    for x in np.concatenate([rng.uniform(1e-300, 1, 200), rng.uniform(1, 1e6, 200), [2.0 ** -54, 1.0, 0.5, 2.0]]):
        assert abs(L.orc_det_log(float(x)) - math.log(x)) <= 4e-16 * max(1.0, abs(math.log(x)))
    for x in np.concatenate([-rng.uniform(0, 700, 300), [0.0, -1e-9, -744.0]]):
        assert abs(L.orc_det_exp(float(x)) - math.exp(x)) <= 1e-15 * math.exp(x) + 5e-324
    assert L.orc_det_exp(-800.0) == 0.0


def test_dirichlet_moments(orc):
    """Dirichlet([a; n]): mean 1/n, variance (n-1)/(n^2 (n a + 1)); no NaN even at alpha = 0.03."""
    for alpha, n, draws in ((0.3, 8, 3000), (0.03, 5, 6000), (1.5, 4, 2000)):
        x = np.stack([orc.dirichlet(9, g, 0, n, alpha) for g in range(draws)]).astype(np.float64)
        assert np.isfinite(x).all() and (x >= 0).all()
        assert np.allclose(x.sum(axis=1), 1.0, atol=1e-5)
        var = (n - 1) / (n * n * (n * alpha + 1))
        se_mean = math.sqrt(var / draws)
        assert np.all(np.abs(x.mean(axis=0) - 1 / n) < 5 * se_mean)
        assert np.all(np.abs(x.var(axis=0) - var) < 0.15 * var)


def test_two_child_alpha003_never_nan(orc):
    """f32 Gamma(0.03) underflows; the log-space draw cannot produce 0/0."""
    x = np.stack([orc.dirichlet(1, g, 7, 2, 0.03) for g in range(4000)])
    assert np.isfinite(x).all() and np.allclose(x.sum(axis=1), 1.0, atol=1e-6)


def test_action_uniform_range(orc):
    u = np.array([orc.lib().orc_action_uniform(5, g, 3) for g in range(2000)])
    assert (u >= 0).all() and (u < 1).all() and abs(u.mean() - 0.5) < 0.03


def test_playout_index_range(orc):
    for n in (1, 2, 15, 126):
        idx = [orc.lib().orc_playout_index(1, g, 0, n) for g in range(500)]
        assert min(idx) >= 0 and max(idx) < n
        if n > 1:
            assert len(set(idx)) > 1
