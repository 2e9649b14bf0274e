// bfb_nuts_dmma_common.cuh -- helpers shared by the tensor-core NUTS kernels (bfb_sampler_dmma.cu, bfb_sampler_pair.cu):
// packed quad reductions of the U-turn tests, one out-of-line instance of the RNG functions and of the dual averaging.
#pragma once
#include "bfb_dmma.cuh"
#include "bfb_nuts_common.cuh"

// per chain: is any of the six sums over its 4 lanes <= 0 ?
__device__ __forceinline__ bool quad_any_nonpos6(double v0, double v1, double v2, double v3, double v4, double v5, int lane)
{
    const bool b0 = lane & 1, b1 = lane & 2;
    double k0 = b0 ? v4 : v0, k1 = b0 ? v5 : v1, k2 = b0 ? 1. : v2, k3 = b0 ? 1. : v3;
    const double s0 = b0 ? v0 : v4, s1 = b0 ? v1 : v5, s2 = b0 ? v2 : 1., s3 = b0 ? v3 : 1.;
    k0 += shx4(s0, 1); k1 += shx4(s1, 1); k2 += shx4(s2, 1); k3 += shx4(s3, 1);
    double m0 = b1 ? k2 : k0, m1 = b1 ? k3 : k1;
    const double t0 = b1 ? k0 : k2, t1 = b1 ? k1 : k3;
    m0 += shx4(t0, 2); m1 += shx4(t1, 2);
    const unsigned bal = __ballot_sync(BFB_FULL, (m0 <= 0.) || (m1 <= 0.));
    return ((bal >> (lane & ~3)) & 0xfu) != 0u;
}

// one instance of Philox + Phi^-1 in the kernel image instead of one per call site (the code of a round must stay
// small: with one or two warps per scheduler instruction-fetch stalls are not hidden by other warps)
static __device__ __noinline__ double draw_normal_ni(uint64_t seed, uint64_t chain, uint64_t t) { return bfb_draw_normal(seed, chain, t); }
static __device__ __noinline__ double draw_uniform_ni(uint64_t seed, uint64_t chain, uint64_t t) { return bfb_draw_uniform(seed, chain, t); }
static __device__ __noinline__ double2 philox_pair_ni(uint64_t seed, uint64_t chain, uint64_t blk)
{
    double u0, u1;
    const bfb_philox_block b = bfb_philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)chain,
                                                 (uint32_t)(chain >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
    u0 = bfb_u64_to_uniform((uint64_t)b.v[0] | ((uint64_t)b.v[1] << 32));
    u1 = bfb_u64_to_uniform((uint64_t)b.v[2] | ((uint64_t)b.v[3] << 32));
    return make_double2(u0, u1);
}
// Nesterov dual averaging of the step size, step_size.py:31-45 -- out of line: pow and exp are ~250 instructions that only
// warm-up iteration boundaries execute, and the hot loop is larger than the instruction cache
static __device__ __noinline__ void dual_average_ni(double cnt, double hbar0, double mu_da, double accept_stat, double log_bar,
                                                    double t0, double target, double gamma, double kk,
                                                    double &hbar, double &log_step, double &log_bar_new, double &e_step, double &e_bar)
{
    const double w = 1. / (cnt + t0);
    hbar = ((1. - w) * hbar0 + w * (target - accept_stat));
    log_step = mu_da - hbar * sqrt(cnt) / gamma;
    const double mk = pow(cnt, -kk);
    log_bar_new = mk * log_step + (1. - mk) * log_bar;
    e_step = exp(log_step); e_bar = exp(log_bar_new);
}
__device__ __forceinline__ int64_t shfl64(int64_t v, int src)
{
    return (int64_t)__shfl_sync(BFB_FULL, (long long)v, src);
}
