// bfb_eval_dmma.cu -- batched surrogate evaluation on the FP64 tensor cores: logp and gradient of C points,
// eight points per warp (bfb_dmma.cuh).  Replaces Density.logp_and_grad (core/density.py:724-754) over
// PolyModel._fun_and_jac (modules/poly.py:443-503) for a surrogate-only density with linear + quadratic (+ cubic-2
// (+ cubic-3)) configs, n <= 32 (28 with cubic-3), with or without radial bound, decay, variable transform and module
// rescale; everything else runs density_eval_kernel (bfb_model.cu).
//
// Persistent blocks of 4 warps; the operand table (29 KB at n = 26 cubic-2) is staged once per block in shared memory and
// every two DMMAs read their B fragments with one conflict-free 16-byte load per lane.  Points are read (software
// pipelined: the next group is in flight during the evaluation) and gradients written in the caller's [C, n] layout: the 4
// lanes of a quad touch one 32-byte sector per r.
#include "bfb_dmma.cuh"
#include <cstring>
#include <cstdlib>

template <int NR, int MV>
__global__ void __launch_bounds__(128, 4) eval_dmma_kernel(DevModel M, const double *__restrict__ X, int64_t C,
                                                        double *__restrict__ LP, double *__restrict__ G)
{
    using SH = DmmaShape<NR, MV>;
    extern __shared__ double bsm[];
    double *msm = bsm + SH::FRAG_DOUBLES;     // per-dimension tables (dmma_stage_tables)
    for (int i = threadIdx.x; i < SH::FRAG_DOUBLES; i += blockDim.x) bsm[i] = M.bfrag[i];
    dmma_stage_tables<MV>(M, msm);
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gi = lane >> 2, lg = lane & 3;
    const int n = M.n;
    const DmmaConsts K = dmma_consts(M);
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
        dmma_logp_grad<NR, MV>(bsm, lane, K, x, msm, valid, lp, gn, [](const double (&)[NR]) { return 0.; }, ke);
        if (valid) {
            if (lg == 0) LP[c] = lp;
#pragma unroll
            for (int r = 0; r < NR; ++r) if (4 * r + lg < n) G[c * n + 4 * r + lg] = gn[r];
        }
    }
}

template <int NR, int MV>
static int launch_eval(bfb_context *h, const double *X, int64_t C, double *LP, double *G)
{
    using SH = DmmaShape<NR, MV>;
    const size_t smem = sizeof(double) * (SH::FRAG_DOUBLES + SH::MSM_DOUBLES + (SH::C3 ? dmma_c3_doubles(NR, h->dm.c3_kt, 4) : 0));
    if (smem > (size_t)(227 * 1024)) return 1;
    BFB_CUDA(cudaFuncSetAttribute(eval_dmma_kernel<NR, MV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 4;
    if (const char *e = getenv("BFB200_EVAL_BLOCKS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 8) per_sm = v; }
    int64_t want = (C + 31) / 32;
    int blocks = (int)(want < (int64_t)h->sm_count * per_sm ? want : (int64_t)h->sm_count * per_sm);
    eval_dmma_kernel<NR, MV><<<blocks, 128, smem, h->stream>>>(h->dm, X, C, LP, G);
    h->launches++;
    BFB_CUDA(cudaGetLastError());
    return BFB_OK;
}

// returns 1 if the tensor-core evaluator does not apply (caller uses the generic kernel), 0 on launch, <0 on error
int bfb_launch_eval_dmma(bfb_context *h, const double *X, int64_t C, double *LP, double *G)
{
    const DevModel &M = h->dm;
    if (M.frag_nr == 0 || (M.has_c3 && (M.c3_kt == 0 || !M.has_c2))) return 1;
    if (const char *e = getenv("BFB200_EVAL")) { if (!strcmp(e, "generic")) return 1; }
    const int mv = (M.has_c2 ? 1 : 0) | (M.frag_ext ? 2 : 0) | (M.has_c3 ? 4 : 0);
#define BFB_CASE(NR_, MV_) if (M.frag_nr == NR_ && mv == MV_) return launch_eval<NR_, MV_>(h, X, C, LP, G);
    BFB_CASE(4, 0) BFB_CASE(4, 1) BFB_CASE(4, 2) BFB_CASE(4, 3) BFB_CASE(7, 0) BFB_CASE(7, 1) BFB_CASE(7, 2) BFB_CASE(7, 3)
    BFB_CASE(8, 0) BFB_CASE(8, 1) BFB_CASE(8, 2) BFB_CASE(8, 3)
    BFB_CASE(4, 5) BFB_CASE(4, 7) BFB_CASE(7, 5) BFB_CASE(7, 7)          // with cubic-3 configs (n <= 28)
#undef BFB_CASE
    return 1;
}
