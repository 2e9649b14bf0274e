// bfb_eval_team.cu -- batched surrogate evaluation for 32 < n <= 64 on the FP64 tensor cores: logp and gradient of C points, eight
// points per TEAM of four warps (bfb_team.cuh: every warp owns 16 of the dimensions, the cubic-3 pair-product operand streams from
// L2).  Replaces Density.logp_and_grad (core/density.py:724-754) over PolyModel._fun_and_jac (modules/poly.py:443-503) for a
// surrogate-only density with linear + quadratic (+ cubic-2 (+ cubic-3)) configs and radial bound (BASELINE configs[3]: 64-D
// cubic-3); below 33 dimensions the one-warp evaluator (bfb_eval_dmma.cu) is the faster one, anything else runs
// density_eval_kernel (bfb_model.cu).  One persistent team per SM: the operand table of [H | S | A | A^T] alone is 128 KB at n = 64.
#include "bfb_team.cuh"
#include <cstring>
#include <cstdlib>

template <int NR, int MV>
__global__ void __launch_bounds__(128, 1) eval_team_kernel(DevModel M, const double *__restrict__ X, int64_t C,
                                                        double *__restrict__ LP, double *__restrict__ G)
{
    using TS = TeamShape<NR, MV>;
    constexpr int NRW = TS::NRW;
    extern __shared__ double smem[];
    double *tab = smem, *msm = smem + TS::TAB_DOUBLES;
    for (int i = threadIdx.x; i < TS::TAB_DOUBLES; i += blockDim.x) tab[i] = M.tfrag[i];
    for (int i = threadIdx.x; i < TS::MSM_HALF && i < M.np; i += blockDim.x) { msm[i] = M.use_bound ? M.mu[i] : 0.; msm[TS::MSM_HALF + i] = M.lin[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gi = lane >> 2, lg = lane & 3;
    double *xbuf = smem + TS::TAB_DOUBLES + TS::MSM, *red = xbuf + 3 * TS::SLOT;
    const double *mu_t = msm;
    const double *tab_w = tab + (size_t)w * NR * TS::NTW * 32;
    const int n = M.n;
    const DmmaConsts K = dmma_consts(M);
    int rbuf = 0;
    const int bar_id = 1;
    double p0[NRW], v1[NRW];
#pragma unroll
    for (int i = 0; i < NRW; ++i) { p0[i] = 0.; v1[i] = 1.; }
    for (int64_t base = (int64_t)blockIdx.x * 8; base < C; base += (int64_t)gridDim.x * 8) {
        const int64_t c = base + gi;
        const bool valid = c < C;
        const int64_t cc = valid ? c : C - 1;
        double q[NRW];
#pragma unroll
        for (int i = 0; i < NRW; ++i) {
            const int j = 4 * (NRW * w + i) + lg, e = (NRW * w + i) * 32 + lane;
            q[i] = (j < n) ? X[cc * n + j] : 0.;
            xbuf[e] = q[i]; xbuf[TS::SLOT + e] = q[i] - mu_t[j]; xbuf[2 * TS::SLOT + e] = q[i] * q[i];
        }
        team_bar(bar_id);
        double gn[NRW], lp, ke;
        team_logp_grad<NR, MV>(tab_w, msm, xbuf, red, rbuf, bar_id, lane, w, K, valid, q, p0, v1, 0., lp, gn, ke);
        if (valid) {
            if (w == 0 && lg == 0) LP[c] = lp;                 // the value is complete in the leader warp only
#pragma unroll
            for (int i = 0; i < NRW; ++i) { const int j = 4 * (NRW * w + i) + lg; if (j < n) G[c * n + j] = gn[i]; }
        }
    }
}

template <int NR, int MV>
static int launch_eval_team(bfb_context *h, const double *X, int64_t C, double *LP, double *G)
{
    using TS = TeamShape<NR, MV>;
    const size_t smem = sizeof(double) * (TS::TAB_DOUBLES + TS::MSM + 3 * TS::SLOT + TS::RED_DOUBLES);
    if (smem > (size_t)(227 * 1024)) return 1;
    BFB_CUDA(cudaFuncSetAttribute(eval_team_kernel<NR, MV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t want = (C + 7) / 8;
    const int blocks = (int)(want < (int64_t)h->sm_count ? want : (int64_t)h->sm_count);
    eval_team_kernel<NR, MV><<<blocks, 128, smem, h->stream>>>(h->dm, X, C, LP, G);
    h->launches++;
    BFB_CUDA(cudaGetLastError());
    return BFB_OK;
}

// returns 1 if this evaluator does not apply (caller tries the next one), 0 on launch, <0 on error
int bfb_launch_eval_team(bfb_context *h, const double *X, int64_t C, double *LP, double *G)
{
    const DevModel &M = h->dm;
    if (M.epilogue || !M.tfrag || M.team_nr != 16 || M.frag_ext) return 1;
    if (M.has_c3 && !(M.tfrag3 && M.has_c2)) return 1;
    if (const char *e = getenv("BFB200_EVAL")) { if (!strcmp(e, "generic")) return 1; }
    if (M.has_c3) return launch_eval_team<16, 5>(h, X, C, LP, G);
    return M.has_c2 ? launch_eval_team<16, 1>(h, X, C, LP, G) : launch_eval_team<16, 0>(h, X, C, LP, G);
}
