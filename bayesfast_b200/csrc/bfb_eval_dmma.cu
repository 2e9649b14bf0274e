// bfb_eval_dmma.cu -- batched surrogate evaluation on the FP64 tensor cores: logp and gradient of C points,
// eight points per warp (bfb_dmma.cuh).  Replaces Density.logp_and_grad (core/density.py:724-754) over
// PolyModel._fun_and_jac (modules/poly.py:443-503) for a surrogate-only density with linear + quadratic (+ cubic-2)
// configs, radial bound, no decay / transform / module rescale; everything else runs density_eval_kernel (bfb_model.cu).
//
// Persistent blocks of 4 warps; the operand table (27 KB at n = 26) is staged once per block in shared memory and
// every DMMA reads its B fragment with one conflict-free 8-byte load per lane.  Points are read and gradients written
// in the caller's [C, n] layout: the 4 lanes of a quad touch one 32-byte sector per r.
#include "bfb_dmma.cuh"
#include <cstring>
#include <cstdlib>

template <int NR, bool C2>
__global__ void __launch_bounds__(128, 4) eval_dmma_kernel(DevModel M, const double *__restrict__ X, int64_t C,
                                                        double *__restrict__ LP, double *__restrict__ G)
{
    using SH = DmmaShape<NR, C2>;
    extern __shared__ double bsm[];
    double *msm = bsm + SH::FRAG_DOUBLES;     // mu[32] | lin[32]
    for (int i = threadIdx.x; i < SH::FRAG_DOUBLES; i += blockDim.x) bsm[i] = M.bfrag[i];
    if (threadIdx.x < 32) {
        msm[threadIdx.x] = M.use_bound ? M.mu[threadIdx.x] : 0.;
        msm[32 + threadIdx.x] = M.lin[threadIdx.x];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gi = lane >> 2, lg = lane & 3;
    const int n = M.n;
    DmmaConsts K;
    K.c0 = M.c0[0]; K.alpha = M.alpha; K.alpha2 = M.use_bound ? M.alpha * M.alpha : INFINITY;
    K.f_mu = M.use_bound ? M.f_mu[0] : 0.; K.n = n;
    const int64_t stride = (int64_t)gridDim.x * 4 * 8;
    int64_t base = ((int64_t)blockIdx.x * 4 + wib) * 8;
    // software pipeline: the points of the next group are in flight while this one is evaluated
    double xn[NR];
    {
        const int64_t cc = (base + gi < C) ? base + gi : C - 1;
#pragma unroll
        for (int r = 0; r < NR; ++r) xn[r] = (4 * r + lg < n && base < C) ? X[cc * n + 4 * r + lg] : 0.;
    }
    for (; base < C; base += stride) {
        const int64_t c = base + gi;
        const bool valid = c < C;
        double x[NR], gn[NR], lp, ke;
#pragma unroll
        for (int r = 0; r < NR; ++r) x[r] = xn[r];
        {
            const int64_t nb = base + stride;
            const int64_t cc = (nb + gi < C) ? nb + gi : C - 1;
#pragma unroll
            for (int r = 0; r < NR; ++r) xn[r] = (4 * r + lg < n && nb < C) ? X[cc * n + 4 * r + lg] : 0.;
        }
        dmma_logp_grad<NR, C2>(bsm, lane, K, x, msm, msm + 32, valid, lp, gn, [](const double (&)[NR]) { return 0.; }, ke);
        if (valid) {
            if (lg == 0) LP[c] = lp;
#pragma unroll
            for (int r = 0; r < NR; ++r) if (4 * r + lg < n) G[c * n + 4 * r + lg] = gn[r];
        }
    }
}

template <int NR, bool C2>
static int launch_eval(bfb_context *h, const double *X, int64_t C, double *LP, double *G)
{
    using SH = DmmaShape<NR, C2>;
    const size_t smem = sizeof(double) * (SH::FRAG_DOUBLES + 64);
    BFB_CUDA(cudaFuncSetAttribute(eval_dmma_kernel<NR, C2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 4;
    if (const char *e = getenv("BFB200_EVAL_BLOCKS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 8) per_sm = v; }
    int64_t want = (C + 31) / 32;
    int blocks = (int)(want < (int64_t)h->sm_count * per_sm ? want : (int64_t)h->sm_count * per_sm);
    eval_dmma_kernel<NR, C2><<<blocks, 128, smem, h->stream>>>(h->dm, X, C, LP, G);
    h->launches++;
    BFB_CUDA(cudaGetLastError());
    return BFB_OK;
}

// returns 1 if the tensor-core evaluator does not apply (caller uses the generic kernel), 0 on launch, <0 on error
int bfb_launch_eval_dmma(bfb_context *h, const double *X, int64_t C, double *LP, double *G)
{
    const DevModel &M = h->dm;
    if (M.frag_nr == 0 || M.has_c3 || M.use_decay || M.use_transform || M.use_scales) return 1;
    if (const char *e = getenv("BFB200_EVAL")) { if (!strcmp(e, "generic")) return 1; }
    const bool c2 = M.has_c2;
    switch (M.frag_nr) {
    case 4: return c2 ? launch_eval<4, true>(h, X, C, LP, G) : launch_eval<4, false>(h, X, C, LP, G);
    case 7: return c2 ? launch_eval<7, true>(h, X, C, LP, G) : launch_eval<7, false>(h, X, C, LP, G);
    case 8: return c2 ? launch_eval<8, true>(h, X, C, LP, G) : launch_eval<8, false>(h, X, C, LP, G);
    }
    return 1;
}
