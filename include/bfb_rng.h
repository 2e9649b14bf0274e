/*
 * bfb_rng.h -- the random-stream contract of bayesfast_b200 (header-only, C99 / CUDA).
 *
 * The reference draws from numpy PCG64 streams, one per chain, obtained with
 * bit_generator.jumped(i + 1) (bayesfast/utils/random.py:20-32, used by
 * bayesfast/samplers/sample_trace.py:199).  That generator cannot be evaluated in
 * lock-step on a GPU, so this build defines its own counter-based stream and feeds
 * the *same numbers* to the oracle when parity is checked.
 *
 * Stream definition (one stream per chain):
 *   draw t (t = 0, 1, 2, ...) of chain c under seed s is
 *       U(s, c, t) = (k + 0.5) * 2^-52,   k = top 52 bits of the 64-bit word
 *       w = philox4x32_10(counter = {lo32(t >> 1), hi32(t >> 1), lo32(c), hi32(c)},
 *                         key     = {lo32(s), hi32(s)})[2 * (t & 1) + {0, 1}]   (word0 | word1 << 32)
 *   so U is strictly inside (0, 1) and exactly representable.
 *   A uniform draw is U itself; a standard normal draw is Phi^-1(U) evaluated with
 *   Wichura's AS 241 PPND16 rational approximation (relative error ~1e-16).
 *   Every draw, normal or uniform, consumes exactly one t.
 *
 * The per-iteration consumption order follows the reference
 * (base_hmc.py:69 -> n normals; nuts.py:210 direction; nuts.py:164 subtree merges in
 * post-order; nuts.py:82 top-level swap; hmc.py:40-41 accept draw): see SURVEY.md 8(a) N-RNG.
 *
 * All floating-point steps use explicit fused multiply-adds so that host and device
 * evaluate the polynomial parts identically; the only library calls are log() and sqrt()
 * in the tails of Phi^-1.
 */
#ifndef BFB_RNG_H
#define BFB_RNG_H

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define BFB_HD __host__ __device__ __forceinline__
#else
#define BFB_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define BFB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define BFB_MULHI32(a, b) __umulhi((a), (b))
#else
#define BFB_FMA(a, b, c) fma((a), (b), (c))
#define BFB_MULHI32(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
#endif

typedef struct { uint32_t v[4]; } bfb_philox_block;

BFB_HD bfb_philox_block bfb_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                          uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = BFB_MULHI32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = BFB_MULHI32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    bfb_philox_block b;
    b.v[0] = c0; b.v[1] = c1; b.v[2] = c2; b.v[3] = c3;
    return b;
}

/* 64-bit word -> U in (0,1): (k + 0.5) * 2^-52 with k the top 52 bits. Exact. */
BFB_HD double bfb_u64_to_uniform(uint64_t w)
{
    return ((double)(w >> 12) + 0.5) * 2.220446049250313e-16; /* 2^-52 */
}

/* draw t of chain c under seed s */
BFB_HD double bfb_draw_uniform(uint64_t seed, uint64_t chain, uint64_t t)
{
    uint64_t blk = t >> 1;
    bfb_philox_block b = bfb_philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32),
                                           (uint32_t)chain, (uint32_t)(chain >> 32),
                                           (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t w = (t & 1) ? ((uint64_t)b.v[2] | ((uint64_t)b.v[3] << 32))
                         : ((uint64_t)b.v[0] | ((uint64_t)b.v[1] << 32));
    return bfb_u64_to_uniform(w);
}

/* Phi^-1(p), Wichura AS 241 (PPND16). p must be in (0,1). */
BFB_HD double bfb_norminv(double p)
{
    double q = p - 0.5, r, num, den, val;
    if (fabs(q) <= 0.425) {
        r = BFB_FMA(-q, q, 0.180625);
        num = 2.5090809287301226727e+3;
        num = BFB_FMA(num, r, 3.3430575583588128105e+4);
        num = BFB_FMA(num, r, 6.7265770927008700853e+4);
        num = BFB_FMA(num, r, 4.5921953931549871457e+4);
        num = BFB_FMA(num, r, 1.3731693765509461125e+4);
        num = BFB_FMA(num, r, 1.9715909503065514427e+3);
        num = BFB_FMA(num, r, 1.3314166789178437745e+2);
        num = BFB_FMA(num, r, 3.3871328727963666080e0);
        den = 5.2264952788528545610e+3;
        den = BFB_FMA(den, r, 2.8729085735721942674e+4);
        den = BFB_FMA(den, r, 3.9307895800092710610e+4);
        den = BFB_FMA(den, r, 2.1213794301586595867e+4);
        den = BFB_FMA(den, r, 5.3941960214247511077e+3);
        den = BFB_FMA(den, r, 6.8718700749205790830e+2);
        den = BFB_FMA(den, r, 4.2313330701600911252e+1);
        den = BFB_FMA(den, r, 1.0);
        return q * num / den;
    }
    r = (q < 0.0) ? p : 1.0 - p;
    r = sqrt(-log(r));
    if (r <= 5.0) {
        r -= 1.6;
        num = 7.74545014278341407640e-4;
        num = BFB_FMA(num, r, 2.27238449892691845833e-2);
        num = BFB_FMA(num, r, 2.41780725177450611770e-1);
        num = BFB_FMA(num, r, 1.27045825245236838258e0);
        num = BFB_FMA(num, r, 3.64784832476320460504e0);
        num = BFB_FMA(num, r, 5.76949722146069140550e0);
        num = BFB_FMA(num, r, 4.63033784615654529590e0);
        num = BFB_FMA(num, r, 1.42343711074968357734e0);
        den = 1.05075007164441684324e-9;
        den = BFB_FMA(den, r, 5.47593808499534494600e-4);
        den = BFB_FMA(den, r, 1.51986665636164571966e-2);
        den = BFB_FMA(den, r, 1.48103976427480074590e-1);
        den = BFB_FMA(den, r, 6.89767334985100004550e-1);
        den = BFB_FMA(den, r, 1.67638483018380384940e0);
        den = BFB_FMA(den, r, 2.05319162663775882187e0);
        den = BFB_FMA(den, r, 1.0);
    } else {
        r -= 5.0;
        num = 2.01033439929228813265e-7;
        num = BFB_FMA(num, r, 2.71155556874348757815e-5);
        num = BFB_FMA(num, r, 1.24266094738807843860e-3);
        num = BFB_FMA(num, r, 2.65321895265761230930e-2);
        num = BFB_FMA(num, r, 2.96560571828504891230e-1);
        num = BFB_FMA(num, r, 1.78482653991729133580e0);
        num = BFB_FMA(num, r, 5.46378491116411436990e0);
        num = BFB_FMA(num, r, 6.65790464350110377720e0);
        den = 2.04426310338993978564e-15;
        den = BFB_FMA(den, r, 1.42151175831644588870e-7);
        den = BFB_FMA(den, r, 1.84631831751005468180e-5);
        den = BFB_FMA(den, r, 7.86869131145613259100e-4);
        den = BFB_FMA(den, r, 1.48753612908506148525e-2);
        den = BFB_FMA(den, r, 1.36929880922735805310e-1);
        den = BFB_FMA(den, r, 5.99832206555887937690e-1);
        den = BFB_FMA(den, r, 1.0);
    }
    val = num / den;
    return (q < 0.0) ? -val : val;
}

BFB_HD double bfb_draw_normal(uint64_t seed, uint64_t chain, uint64_t t)
{
    return bfb_norminv(bfb_draw_uniform(seed, chain, t));
}

#endif /* BFB_RNG_H */
