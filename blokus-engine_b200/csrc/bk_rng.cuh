// bk_rng.cuh — the seeded generator the kernels use where the reference calls rand::thread_rng()
// (self_play/src/simulation.rs:107-109 Dirichlet, :120 gen_range(0.0..1.0), :250 gen_range(0..n)).
//
// rand 0.8.5 / rand_distr 0.4.3 are crates.io dependencies that are not under /root/reference and
// thread_rng cannot be seeded, so the reference's draws are not reproducible; this is a
// counter-based replacement (Philox4x32-10) keyed by (seed, GLOBAL game id, ply, purpose, index):
// results do not depend on how games are sharded over GPUs.  The Dirichlet draw follows
// rand_distr's published method (normalised Gamma(alpha,1); Marsaglia-Tsang for shape >= 1 and
// Gamma(alpha+1)*U^(1/alpha) below 1) in f64 and in log space, because at alpha = 0.03 an f32
// Gamma draw underflows to 0 in ~4.5% of cases and 0/0 priors would follow.
// log/exp are polynomial and use only IEEE-rounded + - * / (explicit __d*_rn, no FMA contraction),
// so the CPU oracle's independent implementation (oracle/rng_oracle.hpp) matches bit for bit.
#pragma once
#include <stdint.h>

enum { BK_RNG_PLAYOUT = 0, BK_RNG_NOISE = 1, BK_RNG_ACTION = 2 };

struct BkPhilox { uint32_t v[4]; };

__device__ __forceinline__ BkPhilox bk_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
    uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, x0), l0 = 0xD2511F53u * x0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, x2), l1 = 0xCD9E8D57u * x2;
        const uint32_t n0 = h1 ^ x1 ^ k0, n2 = h0 ^ x3 ^ k1;
        x0 = n0; x1 = l1; x2 = n2; x3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    BkPhilox o;
    o.v[0] = x0; o.v[1] = x1; o.v[2] = x2; o.v[3] = x3;
    return o;
}

// index into n ascending legal tiles for the random-playout policy: word (ply & 3) of the block keyed by
// ply >> 2, so a persistent kernel runs Philox once per four plies
struct BkPlayoutRng {
    uint32_t w0, w1, w2, w3;
    uint32_t block;   // ply >> 2 the words belong to; 0xffffffff = none
};
__device__ __forceinline__ uint32_t bk_playout_index(uint64_t seed, uint32_t game, uint32_t ply, uint32_t n,
                                                     BkPlayoutRng& rng) {
    if (rng.block != (ply >> 2)) {
        const BkPhilox b = bk_philox(seed, game, ply >> 2, BK_RNG_PLAYOUT, 0u);
        rng.w0 = b.v[0]; rng.w1 = b.v[1]; rng.w2 = b.v[2]; rng.w3 = b.v[3];
        rng.block = ply >> 2;
    }
    const uint32_t k = ply & 3u;
    const uint32_t x = k == 0u ? rng.w0 : (k == 1u ? rng.w1 : (k == 2u ? rng.w2 : rng.w3));
    return __umulhi(x, n);
}

// u in [0,1) for softmax_sample (simulation.rs:120), exact in f32
__device__ __forceinline__ float bk_action_uniform(uint64_t seed, uint32_t game, uint32_t ply) {
    const BkPhilox b = bk_philox(seed, game, ply, BK_RNG_ACTION, 0u);
    return __fmul_rn(float(b.v[0] >> 8), 1.0f / 16777216.0f);
}

__device__ __forceinline__ double bk_open01(uint32_t hi, uint32_t lo) {
    const uint64_t m = (uint64_t(hi) << 20) | (uint64_t(lo) >> 12);
    return __dmul_rn(__dadd_rn(double(m), 0.5), 1.0 / 4503599627370496.0);
}

#define BK_LN2_HI 6.93147180369123816490e-01
#define BK_LN2_LO 1.90821492927058770002e-10

__device__ __forceinline__ double bk_det_log(double x) {
    const uint64_t b = uint64_t(__double_as_longlong(x));
    int e = int(b >> 52) - 1023;
    double m = __longlong_as_double((long long)((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull));
    if (m > 1.4142135623730951) { m = __dmul_rn(m, 0.5); e += 1; }
    const double s = __ddiv_rn(__dadd_rn(m, -1.0), __dadd_rn(m, 1.0));
    const double z = __dmul_rn(s, s);
    double p = 1.0 / 23.0;
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 21.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 19.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 17.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 15.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 13.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 11.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 9.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 7.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 5.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 3.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0);
    const double logm = __dmul_rn(__dmul_rn(2.0, s), p);
    const double de = double(e);
    return __dadd_rn(__dmul_rn(de, BK_LN2_HI), __dadd_rn(__dmul_rn(de, BK_LN2_LO), logm));
}

__device__ __forceinline__ double bk_det_exp(double x) {
    if (x < -745.0) return 0.0;
    const double k = floor(__dadd_rn(__dmul_rn(x, 1.44269504088896338700e+00), 0.5));
    const double r = __dadd_rn(__dadd_rn(x, -__dmul_rn(k, BK_LN2_HI)), -__dmul_rn(k, BK_LN2_LO));
    double p = 1.0 / 6227020800.0;
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 479001600.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 39916800.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 3628800.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 362880.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 40320.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 5040.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 720.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 120.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 24.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0 / 6.0);
    p = __dadd_rn(__dmul_rn(p, r), 0.5);
    p = __dadd_rn(__dmul_rn(p, r), 1.0);
    p = __dadd_rn(__dmul_rn(p, r), 1.0);
    const int ki = int(k);
    if (ki >= -1000) return __dmul_rn(p, __longlong_as_double((long long)(uint64_t(ki + 1023) << 52)));
    return __dmul_rn(__dmul_rn(p, __longlong_as_double((long long)(uint64_t(ki + 1000 + 1023) << 52))),
                     __longlong_as_double((long long)(uint64_t(23) << 52)));
}

// log of one Gamma(alpha,1) draw for root child i of (game, ply)
__device__ __forceinline__ double bk_log_gamma_draw(uint64_t seed, uint32_t game, uint32_t ply, uint32_t i, double alpha) {
    const bool small = alpha < 1.0;
    const double shape = small ? __dadd_rn(alpha, 1.0) : alpha;
    const double d = __dadd_rn(shape, -(1.0 / 3.0));
    const double c = __ddiv_rn(1.0, __dsqrt_rn(__dmul_rn(9.0, d)));
#pragma unroll 1
    for (uint32_t k = 0; k < 32768u; ++k) {
        const BkPhilox a = bk_philox(seed, game, ply, BK_RNG_NOISE, (i << 16) | (2u * k));
        const double g1 = __dadd_rn(__dmul_rn(2.0, bk_open01(a.v[0], a.v[1])), -1.0);
        const double g2 = __dadd_rn(__dmul_rn(2.0, bk_open01(a.v[2], a.v[3])), -1.0);
        const double s = __dadd_rn(__dmul_rn(g1, g1), __dmul_rn(g2, g2));
        if (s >= 1.0 || s == 0.0) continue;
        const double x = __dmul_rn(g1, __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, bk_det_log(s)), s)));
        const double v_cbrt = __dadd_rn(1.0, __dmul_rn(c, x));
        if (v_cbrt <= 0.0) continue;
        const double v = __dmul_rn(__dmul_rn(v_cbrt, v_cbrt), v_cbrt);
        const BkPhilox b = bk_philox(seed, game, ply, BK_RNG_NOISE, (i << 16) | (2u * k + 1u));
        const double u = bk_open01(b.v[0], b.v[1]);
        const double x2 = __dmul_rn(x, x);
        const double lv = bk_det_log(v);
        const bool squeeze = u < __dadd_rn(1.0, -__dmul_rn(__dmul_rn(0.0331, x2), x2));
        if (squeeze || bk_det_log(u) < __dadd_rn(__dmul_rn(0.5, x2),
                                                 __dmul_rn(d, __dadd_rn(__dadd_rn(1.0, -v), lv)))) {
            double lg = __dadd_rn(bk_det_log(d), lv);
            if (small) lg = __dadd_rn(lg, __ddiv_rn(bk_det_log(bk_open01(b.v[2], b.v[3])), alpha));
            return lg;
        }
    }
    return 0.0;
}
