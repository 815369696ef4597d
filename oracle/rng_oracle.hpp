// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product.
//
// The reference draws from rand 0.8.5 `thread_rng()` (OS-seeded ChaCha12; cannot be seeded) and
// rand_distr 0.4.3 `Dirichlet` (self_play/src/simulation.rs:107-109,120,250; Cargo.lock:968,998),
// neither of which is under /root/reference.  "Parity unpinned" for every random draw: there is no
// seed that reproduces a reference run.  This file is an independent host-only implementation of
// the seeded generator SPEC that the B200 kernels implement separately
// (blokus-engine_b200/csrc/bk_rng.cuh); the parity tests prove the two agree bit for bit, and
// tests/test_rng_spec.py checks the Dirichlet draw distributionally (mean, variance).
//
// SPEC
//   block(seed, game, ply, purpose, index) = Philox4x32-10(key = (seed_lo, seed_hi),
//                                                          ctr = (game, ply, purpose, index))
//   PLAYOUT (purpose 0): x = block(seed, game, ply >> 2, PLAYOUT, 0)[ply & 3]  (one block serves four plies);
//                        child index = (u64(x) * n) >> 32 over ascending legal tiles
//   ACTION  (purpose 2): u = f32(block(...,0)[0] >> 8) * 2^-24   (simulation.rs:120 stand-in)
//   NOISE   (purpose 1): Dirichlet([alpha; n]) as normalised Gamma(alpha,1) draws, rand_distr's
//     published method (alpha<1: Gamma(alpha+1)*U^(1/alpha); Marsaglia-Tsang for shape>=1) done in
//     f64 and in LOG space so alpha=0.03 cannot underflow to 0/0:
//       child i, attempt k uses blocks index=(i<<16 | 2k) -> (u1,u2) and (i<<16 | 2k+1) -> (u3,u4)
//       polar normal from (u1,u2), MT accept test with u3, lg = log(d*v) + log(u4)/alpha
//       noise_i = f32( exp(lg_i - max lg) / sum_j exp(lg_j - max lg) )   (sum sequential, ascending)
//     log/exp are the deterministic polynomial versions below (only IEEE + - * / sqrt, no FMA),
//     so a CPU and a GPU evaluate them to identical bits.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

enum RngPurpose : uint32_t { RNG_PLAYOUT = 0, RNG_NOISE = 1, RNG_ACTION = 2 };

struct Philox4 { uint32_t v[4]; };

inline Philox4 philox4x32_10(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
    uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = uint64_t(0xD2511F53u) * x0;
        uint64_t p1 = uint64_t(0xCD9E8D57u) * x2;
        uint32_t n0 = uint32_t(p1 >> 32) ^ x1 ^ k0;
        uint32_t n1 = uint32_t(p1);
        uint32_t n2 = uint32_t(p0 >> 32) ^ x3 ^ k1;
        uint32_t n3 = uint32_t(p0);
        x0 = n0; x1 = n1; x2 = n2; x3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{{x0, x1, x2, x3}};
}

inline uint32_t playout_index(uint64_t seed, uint32_t game, uint32_t ply, uint32_t n) {
    Philox4 b = philox4x32_10(seed, game, ply >> 2, RNG_PLAYOUT, 0);
    return uint32_t((uint64_t(b.v[ply & 3]) * n) >> 32);
}

inline float action_uniform(uint64_t seed, uint32_t game, uint32_t ply) {
    Philox4 b = philox4x32_10(seed, game, ply, RNG_ACTION, 0);
    return float(b.v[0] >> 8) * (1.0f / 16777216.0f);
}

inline double bits_to_double(uint64_t b) { double d; std::memcpy(&d, &b, 8); return d; }
inline uint64_t double_to_bits(double d) { uint64_t b; std::memcpy(&b, &d, 8); return b; }

// uniform in the open interval (0,1): 52 random bits + one half, exact in f64
inline double open01(uint32_t hi, uint32_t lo) {
    uint64_t m = (uint64_t(hi) << 20) | (uint64_t(lo) >> 12);
    return (double(m) + 0.5) * (1.0 / 4503599627370496.0);
}

// log(x) for positive normal x: 2*atanh((m-1)/(m+1)) series, 12 terms, plain * and + only.
inline double det_log(double x) {
    uint64_t b = double_to_bits(x);
    int e = int(b >> 52) - 1023;
    double m = bits_to_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 23.0;
    p = p * z + 1.0 / 21.0;
    p = p * z + 1.0 / 19.0;
    p = p * z + 1.0 / 17.0;
    p = p * z + 1.0 / 15.0;
    p = p * z + 1.0 / 13.0;
    p = p * z + 1.0 / 11.0;
    p = p * z + 1.0 / 9.0;
    p = p * z + 1.0 / 7.0;
    p = p * z + 1.0 / 5.0;
    p = p * z + 1.0 / 3.0;
    p = p * z + 1.0;
    double logm = 2.0 * s * p;
    double de = double(e);
    return de * 6.93147180369123816490e-01 + (de * 1.90821492927058770002e-10 + logm);
}

// exp(x) for x <= 0: k = floor(x/ln2 + 1/2), Taylor degree 13 on the remainder, scale by 2^k.
inline double det_exp(double x) {
    if (x < -745.0) return 0.0;
    double k = std::floor(x * 1.44269504088896338700e+00 + 0.5);
    double r = (x - k * 6.93147180369123816490e-01) - k * 1.90821492927058770002e-10;
    double p = 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    int ki = int(k);
    if (ki >= -1000) return p * bits_to_double(uint64_t(ki + 1023) << 52);
    return (p * bits_to_double(uint64_t(ki + 1000 + 1023) << 52)) * bits_to_double(uint64_t(23) << 52);
}

// log of one Gamma(alpha,1) draw for child i (see SPEC).  rand_distr 0.4.3 gamma.rs published
// algorithm: GammaLargeShape (Marsaglia & Tsang 2000) and GammaSmallShape.
inline double log_gamma_draw(uint64_t seed, uint32_t game, uint32_t ply, uint32_t i, double alpha) {
    const bool small = alpha < 1.0;
    const double shape = small ? alpha + 1.0 : alpha;
    const double d = shape - 1.0 / 3.0;
    const double c = 1.0 / std::sqrt(9.0 * d);
    for (uint32_t k = 0; k < 32768; ++k) {
        Philox4 a = philox4x32_10(seed, game, ply, RNG_NOISE, (i << 16) | (2 * k));
        double g1 = 2.0 * open01(a.v[0], a.v[1]) - 1.0;
        double g2 = 2.0 * open01(a.v[2], a.v[3]) - 1.0;
        double s = g1 * g1 + g2 * g2;
        if (s >= 1.0 || s == 0.0) continue;
        double x = g1 * std::sqrt(-2.0 * det_log(s) / s);
        double v_cbrt = 1.0 + c * x;
        if (v_cbrt <= 0.0) continue;
        double v = v_cbrt * v_cbrt * v_cbrt;
        Philox4 b = philox4x32_10(seed, game, ply, RNG_NOISE, (i << 16) | (2 * k + 1));
        double u = open01(b.v[0], b.v[1]);
        double x2 = x * x;
        double lv = det_log(v);
        if (u < 1.0 - 0.0331 * x2 * x2 || det_log(u) < 0.5 * x2 + d * (1.0 - v + lv)) {
            double lg = det_log(d) + lv;
            if (small) lg = lg + det_log(open01(b.v[2], b.v[3])) / alpha;
            return lg;
        }
    }
    return 0.0;  // unreachable in practice (acceptance ~0.75 per attempt)
}

// Dirichlet([alpha; n]) noise for the root of (game, ply), f32, ascending child order.
inline std::vector<float> dirichlet_noise(uint64_t seed, uint32_t game, uint32_t ply, uint32_t n,
                                          float alpha) {
    std::vector<double> lg(n);
    double mx = 0.0;
    for (uint32_t i = 0; i < n; ++i) {
        lg[i] = log_gamma_draw(seed, game, ply, i, double(alpha));
        if (i == 0 || lg[i] > mx) mx = lg[i];
    }
    std::vector<double> e(n);
    double sum = 0.0;
    for (uint32_t i = 0; i < n; ++i) { e[i] = det_exp(lg[i] - mx); sum = sum + e[i]; }
    std::vector<float> out(n);
    for (uint32_t i = 0; i < n; ++i) out[i] = float(e[i] / sum);
    return out;
}

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace orc
