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

// Feature form (model variant bit 4 of bfb_dmma.cuh): F = Phi(x) C^T and grad = -(dPhi/dx)^T (C^T f) as two chained GEMMs per
// record of 8 outputs.  One group of 8 points per warp; the 4 warps of a block walk over the records in step, chunks of records
// alternate between two shared-memory buffers filled with cp.async.
template <int NR, bool EXT>
__global__ void __launch_bounds__(128) likf_eval_dmma_kernel(const double *__restrict__ ftab, const int *__restrict__ fpt,
                                                             const int *__restrict__ fgt, const double *__restrict__ dense_tab,
                                                             int m, int n, int kt1, int nt2, int nt1, int frec, double e_c0, int chunk,
                                                             const double *__restrict__ X, int64_t C,
                                                             double *__restrict__ LP, double *__restrict__ G, LikExt E)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gi = lane >> 2, lg = lane & 3;
    const double *hrec = dense_tab + (size_t)m * lik_rec_doubles(NR);
    double *et = sm + (size_t)2 * chunk * frec;                       // per-dimension tables of the extended pipeline
    double *xs = et + 288 + (size_t)wib * (256 + 8 * 8 * LIKF_NT2), *ws = xs + 256;
    if (EXT) {
        for (int i = threadIdx.x; i < 288; i += blockDim.x) et[i] = hrec[lik_rec_doubles(NR) + i];
        __syncthreads();
    }
    for (int64_t base = (int64_t)blockIdx.x * 32; base < C; base += (int64_t)gridDim.x * 32) {
        const int64_t c = base + wib * 8 + gi, cc = c < C ? c : C - 1;
        double xt[NR], xe[NR], gr[NR], beta = 0.;
        bool outside = false;
#pragma unroll
        for (int r = 0; r < NR; ++r) { xt[r] = (4 * r + lg < n) ? X[cc * n + 4 * r + lg] : 0.; xe[r] = xt[r]; }
        if (EXT) lik_pre<NR>(et, E, n, lane, xt, hrec, true, xe, outside, beta);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; ++r) xs[gi * 32 + 4 * r + lg] = (r < NR) ? xe[r < NR ? r : 0] : 0.;
        __syncwarp();
        double phi[LIKF_KT];
        likf_features(fpt, kt1, lane, xs, phi);
        LikFeatAcc A;
        A.acc2 = 0.; A.S1 = 0.;
#pragma unroll
        for (int t2 = 0; t2 < LIKF_NT2; ++t2) A.w[t2][0] = A.w[t2][1] = 0.;
        const int n_chunks = (nt1 + chunk - 1) / chunk;
        auto stage = [&](int k) {
            const int t0 = k * chunk, cnt = (nt1 - t0 < chunk) ? nt1 - t0 : chunk;
            const double *src = ftab + (size_t)t0 * frec;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + (size_t)(k & 1) * chunk * frec);
            for (int i = threadIdx.x; i < cnt * (frec / 2); i += blockDim.x)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 16u * i), "l"(src + 2 * i) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        __syncthreads();
        stage(0);
        for (int k = 0; k < n_chunks; ++k) {
            const int t0 = k * chunk, cnt = (nt1 - t0 < chunk) ? nt1 - t0 : chunk;
            if (k + 1 < n_chunks) { stage(k + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            const double *cbase = sm + (size_t)(k & 1) * chunk * frec;
#pragma unroll 1
            for (int tt = 0; tt < cnt; ++tt)
                likf_tile(cbase + (size_t)tt * frec, kt1, nt2, lane, phi, EXT && E.use_bound, outside, beta, E.alpha, A);
            __syncthreads();
        }
        double acc2, S1, lpv;
        likf_finish<NR>(fgt, nt2, n, lane, xs, ws, A, acc2, S1, gr);
        lpv = e_c0 - 0.5 * acc2;
        if constexpr (EXT) lik_post<NR>(et, E, n, lane, xt, hrec, outside, beta, S1, acc2, e_c0, gr, lpv);
        if (c < C) {
            if (lg == 0) LP[c] = lpv;
#pragma unroll
            for (int r = 0; r < NR; ++r) if (4 * r + lg < n) G[c * n + 4 * r + lg] = gr[r];
        }
    }
}

// Feature table of the model currently on the device: applies when the union Q of the inputs the quadratic parts touch is small
// enough (P_f = 1 + n + q (q + 1) / 2 <= 4 LIKF_KT) and the two chained GEMMs cost fewer DMMAs than one n x n product per output.
static int build_likf_table(bfb_context *h, const std::vector<double> &S, const std::vector<double> &lin, const std::vector<double> &c0, int nr)
{
    DevModel &M = h->dm;
    const int n = M.n, np = M.np, m = M.m;
    M.lik_ftab = nullptr; M.lik_kt1 = 0;
    if (getenv("BFB200_LIK_DENSE")) return BFB_OK;
    std::vector<int> Q;
    for (int j = 0; j < n; ++j) {
        bool used = false;
        for (int o = 0; o < m && !used; ++o)
            for (int k = 0; k < n && !used; ++k) used = S[((size_t)o * n + k) * np + j] != 0. || S[((size_t)o * n + j) * np + k] != 0.;
        if (used) Q.push_back(j);
    }
    const int q = (int)Q.size();
    const int Pf = 1 + n + q * (q + 1) / 2;
    const int kt1 = (Pf + 3) / 4, nt2 = (Pf + 7) / 8, nt1 = (m + 7) / 8;
    if (q > LIKF_QMAX || kt1 > LIKF_KT || nt2 > LIKF_NT2) return BFB_OK;
    if ((kt1 + 2 * nt2) >= 8 * nr * ((nr + 1) / 2)) return BFB_OK;                 // DMMAs per 8 outputs: feature form vs one n x n product each
    const int frec = likf_rec_doubles(kt1, nt2);
    // feature list: 0 constant | 1 + j linear | pairs (a <= b) of Q in lexicographic order
    std::vector<int> fpt((size_t)4 * kt1 * 2, -2), fgt((size_t)32 * (1 + 3 * LIKF_QMAX), 0);
    std::vector<std::pair<int, int>> feat;
    feat.push_back({-1, -1});
    for (int j = 0; j < n; ++j) feat.push_back({j, -1});
    for (int a = 0; a < q; ++a)
        for (int b = a; b < q; ++b) {
            const int fi = (int)feat.size(), ja = Q[a], jb = Q[b];
            feat.push_back({ja, jb});
            // d(x_a x_b)/dx_a = x_b (2 x_a if a == b): entries (feature, partner, factor) of both dimensions
            int *ga = fgt.data() + (size_t)ja * (1 + 3 * LIKF_QMAX);
            ga[1 + 3 * ga[0]] = fi; ga[2 + 3 * ga[0]] = jb; ga[3 + 3 * ga[0]] = (ja == jb) ? 2 : 1; ga[0]++;
            if (ja != jb) {
                int *gb = fgt.data() + (size_t)jb * (1 + 3 * LIKF_QMAX);
                gb[1 + 3 * gb[0]] = fi; gb[2 + 3 * gb[0]] = ja; gb[3 + 3 * gb[0]] = 1; gb[0]++;
            }
        }
    for (int f = 0; f < Pf; ++f) { fpt[2 * f] = feat[f].first; fpt[2 * f + 1] = feat[f].second; }
    auto coef = [&](int o, int f) -> double {            // C[o][f]
        if (o >= m || f >= Pf) return 0.;
        const int a = feat[f].first, b = feat[f].second;
        if (a < 0) return c0[o];
        if (b < 0) return lin[(size_t)o * np + a];
        const double sab = S[((size_t)o * n + a) * np + b];      // symmetrised: S[a][b] = a_ab (a < b), 2 a_aa
        return (a == b) ? 0.5 * sab : sab;
    };
    std::vector<double> tab((size_t)nt1 * frec, 0.);
    for (int t = 0; t < nt1; ++t) {
        double *rec = tab.data() + (size_t)t * frec;
        for (int lane = 0; lane < 32; ++lane) {
            const int kk = lane & 3, ncol = lane >> 2;
            const int o1 = 8 * t + 4 * (ncol & 1) + (ncol >> 1);           // first GEMM: column 2 lg + e <-> output 8 t + 4 e + lg
            for (int kt = 0; kt < kt1; ++kt) rec[kt * 32 + lane] = coef(o1, 4 * kt + kk);
            for (int e = 0; e < 2; ++e)
                for (int t2 = 0; t2 < nt2; ++t2) rec[(kt1 + e * nt2 + t2) * 32 + lane] = coef(8 * t + 4 * e + kk, 8 * t2 + ncol);
        }
        for (int i = 0; i < 8; ++i)
            rec[(kt1 + 2 * nt2) * 32 + i] = (M.use_bound && 8 * t + i < m && 8 * t + i < (int)h->h_fmu.size()) ? h->h_fmu[8 * t + i] : 0.;
    }
    void *p = nullptr;
    BFB_CUDA(cudaMalloc(&p, sizeof(double) * tab.size()));
    h->model_allocs.push_back(p);
    BFB_CUDA(cudaMemcpy(p, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
    M.lik_ftab = (const double *)p;
    BFB_CUDA(cudaMalloc(&p, sizeof(int) * fpt.size()));
    h->model_allocs.push_back(p);
    BFB_CUDA(cudaMemcpy(p, fpt.data(), sizeof(int) * fpt.size(), cudaMemcpyHostToDevice));
    M.lik_fpt = (const int *)p;
    BFB_CUDA(cudaMalloc(&p, sizeof(int) * fgt.size()));
    h->model_allocs.push_back(p);
    BFB_CUDA(cudaMemcpy(p, fgt.data(), sizeof(int) * fgt.size(), cudaMemcpyHostToDevice));
    M.lik_fgt = (const int *)p;
    M.lik_kt1 = kt1; M.lik_nt2 = nt2; M.lik_nt1 = nt1; M.lik_frec = frec;
    return BFB_OK;
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
    { int rcf = build_likf_table(h, S, lin, c0, nr); if (rcf) return rcf; }
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
    if (M.lik_ftab) {
        int fchunk = (40 * 1024) / (int)(M.lik_frec * sizeof(double));
        if (fchunk < 1) fchunk = 1;
        if (fchunk > M.lik_nt1) fchunk = M.lik_nt1;
        const size_t fsmem = sizeof(double) * (2 * (size_t)fchunk * M.lik_frec + 288 + 4 * (256 + 8 * 8 * LIKF_NT2));
        const int64_t fwant = (C + 31) / 32, fcap = (int64_t)h->sm_count * 2;
        const int fblocks = (int)(fwant < fcap ? fwant : fcap);
        if (M.lik_ext) {
            BFB_CUDA(cudaFuncSetAttribute(likf_eval_dmma_kernel<NR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            likf_eval_dmma_kernel<NR, true><<<fblocks, 128, fsmem, h->stream>>>(M.lik_ftab, M.lik_fpt, M.lik_fgt, h->lik_tab, M.m, M.n, M.lik_kt1,
                                                                                M.lik_nt2, M.lik_nt1, M.lik_frec, M.e_c0, fchunk, X, C, LP, G, E);
        } else {
            BFB_CUDA(cudaFuncSetAttribute(likf_eval_dmma_kernel<NR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            likf_eval_dmma_kernel<NR, false><<<fblocks, 128, fsmem, h->stream>>>(M.lik_ftab, M.lik_fpt, M.lik_fgt, h->lik_tab, M.m, M.n, M.lik_kt1,
                                                                                 M.lik_nt2, M.lik_nt1, M.lik_frec, M.e_c0, fchunk, X, C, LP, G, E);
        }
    } else if (M.lik_ext) {
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
