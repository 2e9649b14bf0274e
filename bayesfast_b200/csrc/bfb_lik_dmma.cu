// bfb_lik_dmma.cu -- two-module pipeline on the FP64 tensor cores: surrogate with m (pre-whitened) outputs followed by the
// Gaussian-likelihood module (bfb_set_epilogue; core/density.py:487-566, examples/des-y1-w-cosmosis.ipynb cells 12-18):
//
//     logp(x) = c0 - 1/2 sum_o f_o(x)^2,     grad = - sum_o f_o(x) grad f_o(x),     f_o = c_o + l_o . x + 1/2 x^T S_o x
//
// Over the outputs the matrix-vector products y_o = S_o x are ONE GEMM per point, [1 x n] . [n x (m n)], whose operand is
// shared by all points: eight points are the rows of m8n8k4 DMMAs (ownership of bfb_dmma.cuh: a point belongs to a quad,
// lane lg owns the dimensions 4 r + lg, which is both the A fragment and -- with the columns of S_o ordered on the host --
// the C fragment).  The m n x n operand (3.3 MB at n = 26, m = 457) does not fit shared memory: the warps of a block walk
// over the outputs in step, chunks of outputs are staged in shared memory by the whole block, and every warp evaluates PG
// groups of 8 points against a staged chunk, so the table is read from L2 once per 32 PG points; the chunks alternate between two
// buffers filled with cp.async, the copy of chunk k + 1 running under the DMMAs of chunk k.  y_o never leaves the
// registers: f_o (one quad reduction), sum f_o^2 and the gradient accumulate on the fly.
//
// Applies to linear + quadratic configs, n <= 32.  With a radial bound, module rescale, variable transform or a Gaussian prior on
// the inputs (the DES-Y1 example's density, examples/des-y1-w-cosmosis.ipynb cells 12-18) the EXT instantiation wraps the loop
// over the outputs in lik_pre / lik_post (bfb_dmma.cuh); a decay term, cubic configs or n > 32 run density_eval (bfb_eval.cuh)
// on the generic kernel.
#include "bfb_dmma.cuh"
#include <cstring>
#include <cstdlib>


template <int NR, int PG, bool EXT>
__global__ void __launch_bounds__(128) lik_eval_dmma_kernel(const double *__restrict__ tab, int m, int n, double e_c0, int chunk,
                                                            const double *__restrict__ X, int64_t C,
                                                            double *__restrict__ LP, double *__restrict__ G, LikExt E)
{
    constexpr int NT = (NR + 1) / 2, REC = lik_rec_doubles(NR), OL = NR * NT * 32;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gi = lane >> 2, lg = lane & 3;
    const int64_t per_block = 32 * PG;
    // extended pipeline: per-dimension tables (after the H record) staged behind the two chunk buffers
    const double *hrec = tab + (size_t)m * REC;
    double *et = sm + (size_t)2 * chunk * REC;
    if (EXT) {
        for (int i = threadIdx.x; i < 288; i += blockDim.x) et[i] = hrec[REC + i];
        __syncthreads();
    }
    for (int64_t base = (int64_t)blockIdx.x * per_block; base < C; base += (int64_t)gridDim.x * per_block) {
        double x[PG][NR], gr[PG][NR], acc2[PG], xt[EXT ? PG : 1][NR], S1[PG], beta[PG];
        bool outside[PG];
        int64_t c[PG];
#pragma unroll
        for (int g = 0; g < PG; ++g) {
            c[g] = base + (int64_t)(g * 4 + wib) * 8 + gi;
            const int64_t cc = c[g] < C ? c[g] : C - 1;
            acc2[g] = 0.; S1[g] = 0.; beta[g] = 0.; outside[g] = false;
#pragma unroll
            for (int r = 0; r < NR; ++r) { x[g][r] = (4 * r + lg < n) ? X[cc * n + 4 * r + lg] : 0.; gr[g][r] = 0.; }
            if constexpr (EXT) {
#pragma unroll
                for (int r = 0; r < NR; ++r) xt[g][r] = x[g][r];
                lik_pre<NR>(et, E, n, lane, xt[g], hrec, true, x[g], outside[g], beta[g]);
            }
        }
        // chunks of outputs go through two shared-memory buffers: the cp.async copies of chunk k + 1 run under the DMMAs of chunk k
        const int n_chunks = (m + chunk - 1) / chunk;
        auto stage = [&](int k) {
            const int o0 = k * chunk, cnt = (m - o0 < chunk) ? m - o0 : chunk;
            const double *src = tab + (size_t)o0 * REC;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + (size_t)(k & 1) * chunk * REC);
            for (int i = threadIdx.x; i < cnt * (REC / 2); i += blockDim.x)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 16u * i), "l"(src + 2 * i) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        __syncthreads();                                       // the buffers of the previous pass have been consumed
        stage(0);
        for (int k = 0; k < n_chunks; ++k) {
            const int o0 = k * chunk, cnt = (m - o0 < chunk) ? m - o0 : chunk;
            if (k + 1 < n_chunks) { stage(k + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                                   // chunk k has landed for every thread
            const double *cbase = sm + (size_t)(k & 1) * chunk * REC;
#pragma unroll 1
            for (int oo = 0; oo < cnt; ++oo) {
                const double *rec = cbase + (size_t)oo * REC;
                double acc[PG][NT][2];
#pragma unroll
                for (int g = 0; g < PG; ++g)
#pragma unroll
                    for (int t = 0; t < NT; ++t) acc[g][t][0] = acc[g][t][1] = 0.;
#pragma unroll
                for (int kt = 0; kt < NR; ++kt)
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        const double b = rec[(kt * NT + t) * 32 + lane];
#pragma unroll
                        for (int g = 0; g < PG; ++g) dmma884(acc[g][t][0], acc[g][t][1], x[g][kt], b);
                    }
                double lin[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) lin[r] = rec[OL + 4 * r + lg];
                const double c0 = rec[OL + 32];
#pragma unroll
                for (int g = 0; g < PG; ++g) {
                    double fpart = 0.;
#pragma unroll
                    for (int r = 0; r < NR; ++r) fpart = fma(fma(0.5, acc[g][r / 2][r % 2], lin[r]), x[g][r], fpart);
                    double f = c0 + qsum(fpart);
                    if (EXT && E.use_bound) {
                        const double fmu = rec[OL + 33], f0 = f;
                        if (outside[g]) f = (beta[g] * f0 - (beta[g] - E.alpha) * fmu) / E.alpha;
                        S1[g] = fma(f, (f0 - fmu) / E.alpha, S1[g]);
                    }
                    acc2[g] = fma(f, f, acc2[g]);
#pragma unroll
                    for (int r = 0; r < NR; ++r) gr[g][r] = fma(-f, lin[r] + acc[g][r / 2][r % 2], gr[g][r]);
                }
            }
            __syncthreads();                                   // buffer k & 1 is refilled by the stage() of the next iteration
        }
#pragma unroll
        for (int g = 0; g < PG; ++g) {
            double lpv = e_c0 - 0.5 * acc2[g];
            if constexpr (EXT) lik_post<NR>(et, E, n, lane, xt[g], hrec, outside[g], beta[g], S1[g], acc2[g], e_c0, gr[g], lpv);
            if (c[g] < C) {
                if (lg == 0) LP[c[g]] = lpv;
#pragma unroll
                for (int r = 0; r < NR; ++r) if (4 * r + lg < n) G[c[g] * n + 4 * r + lg] = gr[g][r];
            }
        }
    }
}

// Operand table of the model currently on the device (D.S [m][n][np] symmetrised quadratic, D.lin [m][np], D.c0 [m]):
// record o = fragments fr[(kt NT + t) 32 + lane] = S_o[k][j] with k = 4 kt + (lane & 3) and the column of the C fragment
// owned by (lane >> 2) mapped to j = 4 (2 t + e) + own like bfb_upload_model does for output 0 | lin_o[32] | c0_o.
int bfb_build_lik_table(bfb_context *h)
{
    DevModel &M = h->dm;
    h->lik_tab = nullptr; h->lik_nr = 0;
    const int n = M.n, np = M.np, m = M.m;
    const int nr = bfb_frag_nr(n);
    if (nr == 0 || np != 32 || !M.has_quad || M.has_c2 || M.has_c3 || M.use_decay) return BFB_OK;
    const int NT = (nr + 1) / 2, REC = lik_rec_doubles(nr), OL = nr * NT * 32;
    // m output records | the H record (radial bound; fragments like an S_o) | per-dimension tables of the extended pipeline
    std::vector<double> S((size_t)m * n * np), lin((size_t)m * np), c0(m), tab((size_t)(m + 1) * REC + 288, 0.);
    BFB_CUDA(cudaMemcpy(S.data(), M.S, sizeof(double) * S.size(), cudaMemcpyDeviceToHost));
    BFB_CUDA(cudaMemcpy(lin.data(), M.lin, sizeof(double) * lin.size(), cudaMemcpyDeviceToHost));
    BFB_CUDA(cudaMemcpy(c0.data(), M.c0, sizeof(double) * c0.size(), cudaMemcpyDeviceToHost));
    for (int o = 0; o < m; ++o) {
        double *rec = tab.data() + (size_t)o * REC;
        const double *So = S.data() + (size_t)o * n * np;
        for (int kt = 0; kt < nr; ++kt)
            for (int t = 0; t < NT; ++t)
                for (int lane = 0; lane < 32; ++lane) {
                    const int k = 4 * kt + (lane & 3), gid = lane >> 2, own = gid >> 1, e = gid & 1;
                    const int v = 2 * t + e, j = 4 * v + own;
                    if (v < nr && k < n && j < n) rec[(kt * NT + t) * 32 + lane] = So[(size_t)k * np + j];
                }
        for (int j = 0; j < n; ++j) rec[OL + j] = lin[(size_t)o * np + j];
        rec[OL + 32] = c0[o];
        if (M.use_bound && o < (int)h->h_fmu.size()) rec[OL + 33] = h->h_fmu[o];
    }
    const bool ext = M.use_bound || M.use_scales || M.use_transform || M.use_prior;
    if (ext) {
        double *rec = tab.data() + (size_t)m * REC, *et = rec + REC;
        if (M.use_bound)
            for (int kt = 0; kt < nr; ++kt)
                for (int t = 0; t < NT; ++t)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int k = 4 * kt + (lane & 3), gid = lane >> 2, own = gid >> 1, e = gid & 1;
                        const int v = 2 * t + e, j = 4 * v + own;
                        if (v < nr && k < n && j < n) rec[(kt * NT + t) * 32 + lane] = h->h_hess[(size_t)j * n + k];
                    }
        for (int j = 0; j < 32; ++j) { et[2 * 32 + j] = 1.; et[4 * 32 + j] = 1.; }
        for (int j = 0; j < n; ++j) {
            if (M.use_bound) et[j] = h->h_mu[j];
            if (M.use_scales) { et[32 + j] = h->h_s0[j]; et[64 + j] = h->h_sdiff[j]; }
            if (M.use_transform) {
                const double lo = h->h_ranges[2 * j], w = h->h_ranges[2 * j + 1] - h->h_ranges[2 * j];
                et[96 + j] = lo; et[128 + j] = w;
                et[160 + j] = (double)((h->h_hb[2 * j] ? 1 : 0) | (h->h_hb[2 * j + 1] ? 2 : 0));
                et[192 + j] = log(fabs(w));
            }
            if (M.use_prior) { et[224 + j] = h->h_pw[j]; et[256 + j] = h->h_pmu[j]; }
        }
    }
    M.lik_ext = ext ? 1 : 0;
    void *p = nullptr;
    BFB_CUDA(cudaMalloc(&p, sizeof(double) * tab.size()));
    h->model_allocs.push_back(p);
    BFB_CUDA(cudaMemcpy(p, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
    h->lik_tab = (double *)p; h->lik_nr = nr;
    M.lik_tab = h->lik_tab; M.lik_nr = nr;          // the sampler kernels read the table through the model (bfb_dmma.cuh, MV bit 3)
    return BFB_OK;
}

template <int NR>
static int launch_lik(bfb_context *h, const double *X, int64_t C, double *LP, double *G)
{
    constexpr int PG = 2, REC = lik_rec_doubles(NR);
    int chunk = (32 * 1024) / (int)(REC * sizeof(double));           // 2 buffers of <= 32 KB of staged records: 3 blocks per SM
    if (const char *e = getenv("BFB200_LIK_CHUNK")) { int v = atoi(e); if (v >= 1 && (size_t)v * REC * sizeof(double) <= 100 * 1024) chunk = v; }
    if (chunk > h->dm.m) chunk = h->dm.m;
    const size_t smem = sizeof(double) * (2 * (size_t)chunk * REC + 288);
    const int64_t want = (C + 32 * PG - 1) / (32 * PG);
    const int64_t cap = (int64_t)h->sm_count * 3;
    const int blocks = (int)(want < cap ? want : cap);
    const DevModel &M = h->dm;
    LikExt E;
    E.use_transform = M.use_transform; E.use_scales = M.use_scales; E.use_bound = M.use_bound; E.use_prior = M.use_prior;
    E.alpha = M.alpha; E.p_c0 = M.p_c0;
    if (M.lik_ext) {
        BFB_CUDA(cudaFuncSetAttribute(lik_eval_dmma_kernel<NR, PG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lik_eval_dmma_kernel<NR, PG, true><<<blocks, 128, smem, h->stream>>>(h->lik_tab, M.m, M.n, M.e_c0, chunk, X, C, LP, G, E);
    } else {
        BFB_CUDA(cudaFuncSetAttribute(lik_eval_dmma_kernel<NR, PG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lik_eval_dmma_kernel<NR, PG, false><<<blocks, 128, smem, h->stream>>>(h->lik_tab, M.m, M.n, M.e_c0, chunk, X, C, LP, G, E);
    }
    h->launches++;
    BFB_CUDA(cudaGetLastError());
    return BFB_OK;
}

// returns 1 if this evaluator does not apply (caller uses the generic kernel), 0 on launch, <0 on error
int bfb_launch_lik_dmma(bfb_context *h, const double *X, int64_t C, double *LP, double *G)
{
    if (!h->dm.epilogue || !h->lik_tab) return 1;
    if (const char *e = getenv("BFB200_EVAL")) { if (!strcmp(e, "generic")) return 1; }
    switch (h->lik_nr) {
    case 4: return launch_lik<4>(h, X, C, LP, G);
    case 7: return launch_lik<7>(h, X, C, LP, G);
    case 8: return launch_lik<8>(h, X, C, LP, G);
    }
    return 1;
}
