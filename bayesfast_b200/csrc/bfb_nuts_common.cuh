// bfb_nuts_common.cuh -- pieces shared by the multi-chain-per-warp NUTS kernels (bfb_sampler_dmma.cu, bfb_sampler_team.cu,
// bfb_sampler_pair.cu): linear-domain multinomial weights, the run-output descriptor and the work-queue initialiser.
#pragma once
#include "bfb_common.cuh"

struct RunOutDevF {
    bfb_run_out o;
    int32_t n_iter;
};

struct WT { double m; int k; };   // weight = m * 2^k, m in [1, 2) (or m == 0)

__device__ __forceinline__ double pow2i(int d)   // 2^d for d in [-1022, 1023], 0 below, +inf above
{
    if (d < -1022) return 0.;
    if (d > 1023) return INFINITY;
    return __longlong_as_double((long long)(d + 1023) << 52);
}
__device__ __forceinline__ WT wt_from_dE(double dE)
{
    // exp(-dE) = 2^y, y = -dE * log2(e)
    const double y = -dE * 1.4426950408889634;
    const double kf = floor(y);
    WT w;
    w.m = exp2(y - kf);
    w.k = (int)kf;
    return w;
}
__device__ __forceinline__ WT wt_add(WT a, WT b)
{
    WT r;
    if (a.k >= b.k) { r.m = fma(b.m, pow2i(b.k - a.k), a.m); r.k = a.k; }
    else { r.m = fma(a.m, pow2i(a.k - b.k), b.m); r.k = b.k; }
    if (r.m >= 2.) { r.m *= 0.5; r.k += 1; }
    return r;
}
// u * a < b
__device__ __forceinline__ bool wt_select(double u, WT a, WT b)
{
    return u * a.m < b.m * pow2i(b.k - a.k);
}
__device__ __forceinline__ double wt_min1(WT w) { return (w.k >= 0) ? 1. : w.m * pow2i(w.k); }

#define BFB_NSLOT 12   // proposal slots per chain: tree proposal + one per pending subtree (<= L - 1) + the current one

// queue[0] = head, queue[1] = tail, queue[2 .. 2 + n_groups) = chunks finished per group, then ring[n_units]: ids of the
// groups whose next chunk can start (initially every group once; a group is appended again when one of its chunks completes)
// head0 > 0: the first head0 units are pre-assigned (warp slot s = blockIdx.x + gridDim.x * warp takes unit s without an atomic),
// which deals the groups evenly over ALL blocks when there are fewer groups than warp slots
static __global__ void queue_init_kernel(int *queue, int n_groups, int n_units, int head0 = 0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { queue[0] = head0; queue[1] = n_groups; }
    if (i < n_groups) queue[2 + i] = 0;
    if (i < n_units) queue[2 + n_groups + i] = (i < n_groups) ? i : -1;
}
