// bfb_dmma.cuh -- PolyModel value + gradient for EIGHT points per warp on the FP64 tensor cores (mma.sync m8n8k4.f64).
//
// All chains of a run share the coefficient tensors, so the four matrix-vector products of a cubic-2 stack
//     y = S x        (quadratic, symmetrised;  _poly.pyx:13-43)
//     t = A x        (cubic-2 row sums;        _poly.pyx:49-64)
//     u = A^T x^2    (cubic-2 column sums;     _poly.pyx:70-80)
//     h = H (x - mu) (Mahalanobis radius of the radial bound; poly.py:466-469)
// are, over the chains, GEMMs  [chains x n] * [n x 4n].  DMMA runs on the same pipe as DFMA (37 TFLOP/s both, measured),
// but one m8n8k4 replaces eight warp-wide DFMAs, so the issue slots, the register-file ports and the shared-memory loads
// that limited the FMA formulations (profiles/r01_a, r01_i: FP64 pipe 16 % busy) are no longer the bound.
//
// Ownership: a point (chain) is owned by the 4 lanes of a quad; lane lg = lane & 3 of the quad owns the dimensions
// j = 4 r + lg, r < NR.  With the points as the M rows of the MMA that is exactly the A fragment (row = lane >> 2,
// column k = 4 kt + lg) and, because the columns of the coefficient operand may be ordered freely on the host, also the
// C fragment: column 2 lg + e of tile tau is assigned the value index v = 2 tau + e of lane lg (dimension r = v, or r =
// v - NR for the second matrix of a block).  No shuffles or shared-memory transposes between x and the results.
//
// Coefficient operand: host-built table bfrag[kt][tile pair][lane][2] (bfb_upload_model, bfb_model.cu), tiles grouped in
// blocks by the A operand they multiply:  [S | A] . x   |   A^T . x^2   |   H . (x - mu).
#pragma once
#include "bfb_common.cuh"
#include "bfb_eval.cuh"      // to_original_1

// MV = model variant: bit 0 = cubic-2 configs present, bit 1 = extended density (decay ellipsoid, variable transform,
// module rescale: core/density.py:724-754, core/module.py:47-96) -- a second H block and the elementwise maps around it,
// bit 2 = cubic-3 configs present: the gradient of P3 = sum_(j<k<l) a_jkl x_j x_k x_l is the GEMM
// [chains x pairs (k<l)] . [pairs x n] of the pair products x_k x_l with T[(k,l)][j] = a_sorted(j,k,l) (0 if j in {k,l});
// its operand table (k-tiles x N3T tiles) follows the per-dimension tables in shared memory, P3 = (sum_j x_j dP3/dx_j) / 3
// bit 3 = likelihood pipeline (bfb_set_epilogue, bfb_lik_dmma.cu): logp = c0 - 1/2 sum_o f_o^2 over the m pre-whitened
// quadratic outputs; the m n x n operand is streamed from L2 (one record per output), nothing is staged in shared memory.
// feature form of the likelihood pipeline: at most LIKF_KT k-tiles of features (P_f <= 4 LIKF_KT) and LIKF_NT2 feature tiles
// of the second GEMM; per warp two staged records of 8 outputs, the point (8 x 32) and the second GEMM's result (8 x 8 LIKF_NT2)
#define LIKF_KT 20
#define LIKF_NT2 10
#define LIKF_QMAX 12
__host__ __device__ constexpr int likf_rec_doubles(int kt1, int nt2) { return (kt1 + 2 * nt2) * 32 + 8; }
#define LIKF_WARP_DOUBLES (2 * likf_rec_doubles(LIKF_KT, LIKF_NT2) + 256 + 8 * 8 * LIKF_NT2)
__host__ __device__ constexpr int lik_rec_doubles(int NR) { return NR * ((NR + 1) / 2) * 32 + 40; }   // fragments | lin[32] | c0 | pad

template <int NR, int MV>
struct DmmaShape {
    static constexpr bool C2 = MV & 1, EXT = (MV & 2) != 0, C3 = (MV & 4) != 0, LIK = (MV & 8) != 0, FEAT = (MV & 16) != 0;
    static constexpr int N3T = (NR + 1) / 2;              // output tiles of the cubic-3 GEMM
    static constexpr int TX = C2 ? NR : (NR + 1) / 2;     // tiles of the block multiplying x
    static constexpr int T2 = C2 ? (NR + 1) / 2 : 0;      // tiles multiplying x^2
    static constexpr int TD = (NR + 1) / 2;               // tiles multiplying x - mu
    static constexpr int TD2 = EXT ? (NR + 1) / 2 : 0;    // tiles multiplying x_orig - mu_decay
    static constexpr int O_D2 = TD, O_X = TD + TD2, O_X2 = O_X + TX;
    static constexpr int NT = TX + T2 + TD + TD2;
    static constexpr int MSM_DOUBLES = EXT ? 288 : 64;    // per-dimension tables staged next to the operand table
    static constexpr int NTP = (NT + 1) / 2;              // tile pairs: the B fragments of two tiles are one 16-byte load
    static constexpr int FRAG_DOUBLES = LIK ? 0 : NR * NTP * 64;
    // shared memory in front of the per-dimension tables: the operand table, or (likelihood pipeline) two staged operand
    // records per warp, filled by the warp itself with cp.async one output ahead of the DMMAs
    __host__ __device__ static constexpr int frag_doubles(int warps) { return FEAT ? warps * LIKF_WARP_DOUBLES : LIK ? warps * 2 * lik_rec_doubles(NR) : FRAG_DOUBLES; }
};

inline int bfb_frag_tiles(int nr, bool c2, bool ext) { return (c2 ? nr : (nr + 1) / 2) + (c2 ? (nr + 1) / 2 : 0) + (ext ? 2 : 1) * ((nr + 1) / 2); }
// instantiated dims-per-lane for input_size n (0: not supported)
// doubles of shared memory after the per-dimension tables for the cubic-3 block: operand table, pair table, per-warp x scratch
__host__ __device__ inline int dmma_c3_doubles(int nr, int c3_kt, int nwarps) { return c3_kt * ((nr + 1) / 2) * 32 + c3_kt * 2 + nwarps * 256; }
inline int bfb_frag_nr(int n) { return n <= 16 ? 4 : n <= 28 ? 7 : n <= 32 ? 8 : 0; }

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double shx4(double v, int m) { return __shfl_xor_sync(BFB_FULL, v, m); }

// four sums over the 4 lanes of a quad at once; every lane of the quad ends with all four totals, bitwise identical
__device__ __forceinline__ void qsum4(double &a, double &b, double &c, double &d, int lane)
{
    const bool b0 = lane & 1, b1 = lane & 2;
    double k0 = b0 ? c : a, k1 = b0 ? d : b;
    const double s0 = b0 ? a : c, s1 = b0 ? b : d;
    k0 += shx4(s0, 1); k1 += shx4(s1, 1);
    double k = b1 ? k1 : k0;
    const double s = b1 ? k0 : k1;
    k += shx4(s, 2);
    const double o2 = shx4(k, 2);
    const double p0 = b1 ? o2 : k, p1 = b1 ? k : o2;
    const double q0 = shx4(p0, 1), q1 = shx4(p1, 1);
    a = b0 ? q0 : p0; b = b0 ? q1 : p1; c = b0 ? p0 : q0; d = b0 ? p1 : q1;
}
__device__ __forceinline__ double qsum(double v)
{
    v += shx4(v, 1);
    v += shx4(v, 2);
    return v;
}

// ----------------------------------------------------------------------------------------------------------------------
// Extended likelihood pipeline (model variant bits 3 | 1): the DES-Y1 example's density shape (examples/des-y1-w-cosmosis.ipynb
// cells 12-18) -- variable transform with hard bounds, module rescale, radial bound, m whitened outputs, Gaussian prior on
// the original-space inputs -- around the per-output DMMAs.  Per point only the scalars of the bound (beta, outside) live
// across the loop over the outputs; everything elementwise is recomputed afterwards, and the bound's Jacobian term is
// applied ONCE to the accumulated gradient:  with G0 = -sum_o f_o J0_o and S1 = sum_o f_o (f0_o - f_mu_o) / alpha,
//     grad = G0 - (S1 + (G0 . d) / beta) (H d / beta)          (PolyModel._fj_bound, poly.py:480-503, summed over the outputs)
// Record m of the operand table holds H (fragments, like an S_o); slot OL + 33 of record o holds f_mu_o.
// Per-dimension tables et[k * 32 + j]: k = 0 mu, 1 s0, 2 sdiff, 3 r_lo, 4 r_w, 5 hard-bound code, 6 log|r_w|, 7 p_w, 8 p_mu.
// ----------------------------------------------------------------------------------------------------------------------
struct LikExt {
    int use_transform, use_scales, use_bound, use_prior;
    double alpha, p_c0;
};

template <int NR>
__device__ __forceinline__ void lik_elementwise(const double *et, const LikExt &E, int n, int lg, const double (&xt)[NR],
                                                double (&xo)[NR], double (&xs)[NR], double (&tj)[NR], double (&tjj)[NR])
{
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int j = 4 * r + lg;
        double v = xt[r];
        tj[r] = 1.; tjj[r] = 0.;
        if (E.use_transform && j < n) {
            const int hbj = (int)et[5 * 32 + j];
            if (hbj != 0) to_original_1(xt[r], et[3 * 32 + j], et[4 * 32 + j], hbj, v, tj[r], tjj[r]);
            else { v = et[3 * 32 + j] + xt[r] * et[4 * 32 + j]; tj[r] = et[4 * 32 + j]; }
        }
        xo[r] = v;
        if (E.use_scales) v = (v - et[1 * 32 + j]) / et[2 * 32 + j];
        xs[r] = (j < n) ? v : 0.;
    }
}

// H d for the 8 points of the warp against the H record (global / L2: two calls per evaluation, not worth staging)
template <int NR>
__device__ __forceinline__ void lik_hd(const double *hrec, int lane, const double (&d)[NR], double (&Hd)[NR])
{
    constexpr int NT4 = (NR + 1) / 2;
    double a_[NT4][2];
#pragma unroll
    for (int t = 0; t < NT4; ++t) a_[t][0] = a_[t][1] = 0.;
#pragma unroll
    for (int kt = 0; kt < NR; ++kt)
#pragma unroll
        for (int t = 0; t < NT4; ++t) dmma884(a_[t][0], a_[t][1], d[kt], __ldg(hrec + (kt * NT4 + t) * 32 + lane));
#pragma unroll
    for (int r = 0; r < NR; ++r) Hd[r] = a_[r / 2][r % 2];
}

// before the loop over the outputs: the point the outputs are evaluated at (x rescaled, or its projection onto the bound)
template <int NR>
__device__ __forceinline__ void lik_pre(const double *et, const LikExt &E, int n, int lane, const double (&xt)[NR], const double *hrec,
                                        bool live, double (&xe)[NR], bool &outside, double &beta)
{
    const int lg = lane & 3;
    double xo[NR], tj[NR], tjj[NR];
    lik_elementwise<NR>(et, E, n, lg, xt, xo, xe, tj, tjj);
    outside = false; beta = 0.;
    if (E.use_bound) {
        double d[NR], Hd[NR], part = 0.;
#pragma unroll
        for (int r = 0; r < NR; ++r) d[r] = (4 * r + lg < n) ? xe[r] - et[4 * r + lg] : 0.;
        lik_hd<NR>(hrec, lane, d, Hd);
#pragma unroll
        for (int r = 0; r < NR; ++r) part = fma(d[r], Hd[r], part);
        beta = sqrt(qsum(part));
        outside = live && (beta > E.alpha);
        if (outside) {
#pragma unroll
            for (int r = 0; r < NR; ++r) xe[r] = (4 * r + lg < n) ? (E.alpha * xe[r] + (beta - E.alpha) * et[4 * r + lg]) / beta : 0.;
        }
    }
}

// after the loop: gr holds G0 = -sum_o f_o J0_o on entry, the gradient with respect to the transformed point on exit
template <int NR>
__device__ __forceinline__ void lik_post(const double *et, const LikExt &E, int n, int lane, const double (&xt)[NR], const double *hrec,
                                         bool outside, double beta, double S1, double acc2, double e_c0, double (&gr)[NR], double &lp)
{
    const int lg = lane & 3;
    double xo[NR], xs[NR], tj[NR], tjj[NR];
    lik_elementwise<NR>(et, E, n, lg, xt, xo, xs, tj, tjj);
    if (E.use_bound && __any_sync(BFB_FULL, outside)) {
        double d[NR], Hd[NR], part = 0.;
#pragma unroll
        for (int r = 0; r < NR; ++r) { d[r] = (4 * r + lg < n) ? xs[r] - et[4 * r + lg] : 0.; part = fma(gr[r], d[r], part); }
        lik_hd<NR>(hrec, lane, d, Hd);
        const double g0d = qsum(part);
        if (outside) {
            const double c = (S1 + g0d / beta) / beta;
#pragma unroll
            for (int r = 0; r < NR; ++r) gr[r] = fma(-c, Hd[r], gr[r]);
        }
    }
    double ppart = 0., tpart = 0.;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int j = 4 * r + lg;
        double g = gr[r];
        if (E.use_scales) g = g / et[2 * 32 + j];
        if (E.use_prior && j < n) {
            const double dk = xo[r] - et[8 * 32 + j], wk = et[7 * 32 + j];
            ppart = fma(wk * dk, dk, ppart);
            g -= wk * dk;
        }
        g *= tj[r];
        if (E.use_transform && j < n) {
            if ((int)et[5 * 32 + j] != 0) { tpart += log(fabs(tj[r])); g += tjj[r] / tj[r]; }
            else tpart += et[6 * 32 + j];
        }
        gr[r] = (j < n) ? g : 0.;
    }
    double z0 = 0., z1 = 0.;
    qsum4(ppart, tpart, z0, z1, lane);
    lp = e_c0 - 0.5 * acc2;
    if (E.use_prior) lp += E.p_c0 - 0.5 * ppart;
    if (E.use_transform) lp += tpart;
}

// ----------------------------------------------------------------------------------------------------------------------
// Feature form of the likelihood pipeline (model variant bit 4).  When the quadratic configs of all outputs live on one small
// set Q of inputs (the DES-Y1 example: one 9-D mask shared by 457 outputs, examples/des-y1-w-cosmosis.ipynb cell 18), the m
// outputs are  F = Phi(x) C^T  over the P_f = 1 + n + q (q + 1) / 2 features  1 | x_j | x_a x_b (a <= b in Q)  and the gradient is
//     grad = - sum_o f_o grad f_o = - (dPhi/dx)^T (C^T f) :
// two chained GEMMs per 8 outputs, [8 points x P_f] . [P_f x 8] and [8 points x 8] . [8 x P_f], 4 m P_f flops per point instead
// of the m (2 n^2 + 5 n) of one n x n product per output (73 features against 27 x 27: 5.7 x less at the DES shape).  The
// columns of an output tile are ordered so that the C fragment of the first GEMM (lane lg: outputs 8 t + lg and 8 t + 4 + lg)
// IS the A fragment of the two k-tiles of the second: the outputs never leave the registers.
// Record of output tile t: g1[kt][lane] (KT1 x 32) | g2[e][t2][lane] (2 x NT2 x 32) | f_mu[8].
// ----------------------------------------------------------------------------------------------------------------------
struct LikFeatAcc {
    double acc2, S1;                    // lane partials: sum f^2, sum f (f0 - f_mu) / alpha over the lane's outputs
    double w[LIKF_NT2][2];              // second GEMM: C^T f, lane lg holds features 8 t2 + 2 lg + {0, 1} of its point
};

// the features of the 8 points of the warp as A fragments: lane lg of quad gi holds Phi[point gi][4 kt + lg]
__device__ __forceinline__ void likf_features(const int *fpt, int kt1, int lane, const double *xs /* [8][32] */, double (&phi)[LIKF_KT])
{
    const double *xr = xs + (lane >> 2) * 32;
#pragma unroll
    for (int kt = 0; kt < LIKF_KT; ++kt) {
        double v = 0.;
        if (kt < kt1) {
            const int2 ab = __ldg(reinterpret_cast<const int2 *>(fpt) + 4 * kt + (lane & 3));
            v = (ab.x == -1) ? 1. : (ab.x < 0) ? 0. : (ab.y < 0) ? xr[ab.x] : xr[ab.x] * xr[ab.y];
        }
        phi[kt] = v;
    }
}

// one record (8 outputs) against the features: both GEMMs, bound extrapolation of the values in between
__device__ __forceinline__ void likf_tile(const double *rec, int kt1, int nt2, int lane, const double (&phi)[LIKF_KT], bool use_bound,
                                          bool outside, double beta, double alpha, LikFeatAcc &A)
{
    double f0 = 0., f1 = 0.;
#pragma unroll
    for (int kt = 0; kt < LIKF_KT; ++kt)
        if (kt < kt1) dmma884(f0, f1, phi[kt], rec[kt * 32 + lane]);
    if (use_bound) {
        const double m0 = rec[(kt1 + 2 * nt2) * 32 + (lane & 3)], m1 = rec[(kt1 + 2 * nt2) * 32 + 4 + (lane & 3)];
        const double g0 = f0, g1 = f1;
        if (outside) { f0 = (beta * g0 - (beta - alpha) * m0) / alpha; f1 = (beta * g1 - (beta - alpha) * m1) / alpha; }
        A.S1 = fma(f0, (g0 - m0) / alpha, A.S1);
        A.S1 = fma(f1, (g1 - m1) / alpha, A.S1);
    }
    A.acc2 = fma(f0, f0, A.acc2);
    A.acc2 = fma(f1, f1, A.acc2);
    const double *g2 = rec + kt1 * 32 + lane;
#pragma unroll
    for (int t2 = 0; t2 < LIKF_NT2; ++t2)
        if (t2 < nt2) {
            dmma884(A.w[t2][0], A.w[t2][1], f0, g2[t2 * 32]);
            dmma884(A.w[t2][0], A.w[t2][1], f1, g2[(nt2 + t2) * 32]);
        }
}

// after the last record: totals over the quad, and G0 = -(dPhi/dx)^T w for the own dimensions (ws: [8][8 LIKF_NT2] scratch)
template <int NR>
__device__ __forceinline__ void likf_finish(const int *fgt, int nt2, int n, int lane, const double *xs, double *ws, LikFeatAcc &A,
                                            double &acc2, double &S1, double (&g0)[NR])
{
    const int gi = lane >> 2, lg = lane & 3;
    double z0 = 0., z1 = 0.;
    acc2 = A.acc2; S1 = A.S1;
    qsum4(acc2, S1, z0, z1, lane);
    double *wr = ws + gi * (8 * LIKF_NT2);
    __syncwarp();
#pragma unroll
    for (int t2 = 0; t2 < LIKF_NT2; ++t2)
        if (t2 < nt2) *reinterpret_cast<double2 *>(wr + 8 * t2 + 2 * lg) = make_double2(A.w[t2][0], A.w[t2][1]);
    __syncwarp();
    const double *xr = xs + gi * 32;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int j = 4 * r + lg;
        const int *gt = fgt + j * (1 + 3 * LIKF_QMAX);
        double s = (j < n) ? wr[1 + j] : 0.;          // linear feature of dimension j
        const int cnt = __ldg(gt);
        for (int i = 0; i < cnt; ++i) {
            const int fi = __ldg(gt + 1 + 3 * i), b = __ldg(gt + 2 + 3 * i), fac = __ldg(gt + 3 + 3 * i);
            s = fma((double)fac * wr[fi], xr[b], s);
        }
        g0[r] = -s;
    }
    __syncwarp();
}

// One evaluation = two GEMM stages.  Stage A: h = H (x - mu) (the TD tiles of the D block; with the extended density also
// h2 = H_decay (x_orig - mu_decay)) gives the Mahalanobis radius, i.e. the inside / outside decision of the radial bound
// (poly.py:466-469).  Outside points are then REPLACED by their projection onto the ellipsoid before stage B, so the
// polynomial blocks ([S | A] . x, A^T . x^2) are evaluated exactly once per point -- at x inside, at x_0 outside, which is all
// PolyModel._fj_bound (poly.py:480-503) needs.  Table tile order: D | D2 | x block (TX tiles) | x^2 block (T2 tiles).
template <int NR, int MV, int T_LO, int T_HI>
__device__ __forceinline__ void dmma_tiles(const double *bsm, int lane, const double (&aD)[NR], const double (&aD2)[NR],
                                           const double (&aX)[NR], const double (&aX2)[NR],
                                           double (&acc)[DmmaShape<NR, MV>::NT][2])
{
    using SH = DmmaShape<NR, MV>;
    const double2 *bp = reinterpret_cast<const double2 *>(bsm) + lane;
    // the table is loop invariant for the callers; without this fence the compiler hoists all B fragments into
    // registers (and spills them) instead of streaming them from shared memory next to the MMAs
    asm volatile("" ::: "memory");
#pragma unroll
    for (int kt = 0; kt < NR; ++kt) {
#pragma unroll
        for (int tp = T_LO / 2; tp < (T_HI + 1) / 2; ++tp) {
            const double2 b = bp[(kt * SH::NTP + tp) * 32];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int t = 2 * tp + e;
                if (t >= T_LO && t < T_HI) {
                    const double a = (t < SH::O_D2) ? aD[kt] : (t < SH::O_X) ? aD2[kt] : (t < SH::O_X2) ? aX[kt] : aX2[kt];
                    dmma884(acc[t][0], acc[t][1], a, e ? b.y : b.x);
                }
            }
        }
    }
}

struct DmmaConsts {
    double c0, alpha, alpha2, f_mu;
    int n;
    // extended density
    double d_alpha2, d_gamma;
    int use_transform, use_scales, use_decay;
    int c3_kt;          // k-tiles of the cubic-3 GEMM (pairs / 4, rounded up)
    // likelihood pipeline
    const double *lik_tab;
    int m;
    double e_c0;
    LikExt lx;
    const double *lik_ftab; const int *lik_fpt, *lik_fgt;
    int kt1, nt2, nt1, frec;
    // cubic-3 block of the team evaluator (bfb_team.cuh): operand streamed from L2
    const double *team3; const int *team_pairs; int team3_kt;
};

__device__ __forceinline__ DmmaConsts dmma_consts(const DevModel &M)
{
    DmmaConsts K;
    K.c0 = M.c0[0]; K.alpha = M.alpha; K.alpha2 = M.use_bound ? M.alpha * M.alpha : INFINITY;
    K.f_mu = M.use_bound ? M.f_mu[0] : 0.; K.n = M.n;
    K.d_alpha2 = M.d_alpha2; K.d_gamma = M.d_gamma;
    K.use_transform = M.use_transform; K.use_scales = M.use_scales; K.use_decay = M.use_decay;
    K.c3_kt = M.c3_kt;
    K.lik_tab = M.lik_tab; K.m = M.m; K.e_c0 = M.e_c0;
    K.lx.use_transform = M.use_transform; K.lx.use_scales = M.use_scales; K.lx.use_bound = M.use_bound; K.lx.use_prior = M.use_prior;
    K.lx.alpha = M.alpha; K.lx.p_c0 = M.p_c0;
    K.lik_ftab = M.lik_ftab; K.lik_fpt = M.lik_fpt; K.lik_fgt = M.lik_fgt;
    K.kt1 = M.lik_kt1; K.nt2 = M.lik_nt2; K.nt1 = M.lik_nt1; K.frec = M.lik_frec;
    K.team3 = M.tfrag3; K.team_pairs = M.tpair; K.team3_kt = M.t3_kt;
    return K;
}

// per-dimension tables in shared memory (first 32 threads of the block): mu | lin  [| s0 | sdiff | d_mu | r_lo | r_w | hb | log|r_w|]
template <int MV>
__device__ __forceinline__ void dmma_stage_tables(const DevModel &M, double *msm)
{
    if ((MV & 8) && (MV & 2)) {
        // extended likelihood pipeline: the per-dimension tables of lik_pre / lik_post
        if (threadIdx.x < 32) {
            const int j = threadIdx.x;
            msm[j] = M.use_bound ? M.mu[j] : 0.;
            msm[32 + j] = M.use_scales ? M.s0[j] : 0.; msm[64 + j] = M.use_scales ? M.sdiff[j] : 1.;
            msm[96 + j] = M.use_transform ? M.r_lo[j] : 0.; msm[128 + j] = M.use_transform ? M.r_w[j] : 1.;
            msm[160 + j] = M.use_transform ? (double)M.hb[j] : 0.;
            msm[192 + j] = M.use_transform ? log(fabs(M.r_w[j])) : 0.;
            msm[224 + j] = M.use_prior ? M.p_w[j] : 0.; msm[256 + j] = M.use_prior ? M.p_mu[j] : 0.;
        }
        return;
    }
    if (threadIdx.x < 32) {
        const int j = threadIdx.x;
        msm[j] = M.use_bound ? M.mu[j] : 0.;
        msm[32 + j] = M.lin[j];
        if (MV & 2) {
            msm[64 + j] = M.s0[j]; msm[96 + j] = M.sdiff[j]; msm[128 + j] = M.d_mu[j];
            msm[160 + j] = M.r_lo[j]; msm[192 + j] = M.r_w[j]; msm[224 + j] = (double)M.hb[j];
            msm[256 + j] = log(fabs(M.r_w[j]));     // log |dx/dx~| of an unbounded (affine) coordinate: constant
        }
    }
    if (MV & 4) {
        // cubic-3 block: operand table [kt][tile][lane], then the pair table (k | l << 8) as ints
        const int nd = M.c3_kt * M.c3_n3t * 32;
        double *t3 = msm + ((MV & 2) ? 288 : 64);
        for (int i = threadIdx.x; i < nd; i += blockDim.x) t3[i] = M.bfrag3[i];
        int *pr = reinterpret_cast<int *>(t3 + nd);
        for (int i = threadIdx.x; i < M.c3_kt * 4; i += blockDim.x) pr[i] = M.c3pair[i];
    }
}

// Density.logp_and_grad(x, original_space=False, use_surrogate=True) of a surrogate-only density for the 8 points of the
// warp: PolyModel._fun_and_jac with the radial bound (poly.py:443-503), and with MV & 2 the variable transform, module
// rescale and decay around it (core/density.py:724-754, core/module.py:80-85; same order of operations as density_eval,
// bfb_eval.cuh).  msm: tables of dmma_stage_tables.  `ke_of` maps the final gradient (own dims) to a lane partial that is
// reduced over the quad (the sampler's kinetic energy of the new momentum); `live` masks points whose outside test must not
// count.  x_in is the point in the sampler's (transformed) space and is left unchanged.
template <int NR, int MV, class KE>
__device__ __forceinline__ void dmma_logp_grad(const double *bsm, int lane, const DmmaConsts &K,
                                               const double (&x_in)[NR], const double *msm,
                                               bool live, double &lp, double (&gn)[NR], KE &&ke_of, double &ke)
{
    using SH = DmmaShape<NR, MV>;
    constexpr bool C2 = SH::C2, EXT = SH::EXT;
    if constexpr (SH::LIK) {
        // likelihood pipeline: y_o = S_o x for every output as DMMAs against the output's record (the 8 points of the warp are
        // the rows), f_o by one quad reduction, sum f_o^2 and the gradient accumulated on the fly (bfb_lik_dmma.cu).  The
        // records stream from L2 through two per-warp shared-memory slots: cp.async of output o + 1 runs under the DMMAs of o.
        if constexpr (SH::FEAT) {
            // feature form: the records (8 outputs each) stream from L2 through two per-warp slots like below
            const int REC = K.frec, gi_ = lane >> 2, lg_ = lane & 3;
            double *wb = const_cast<double *>(bsm) + (size_t)(threadIdx.x >> 5) * LIKF_WARP_DOUBLES;
            double *xs = wb + 2 * likf_rec_doubles(LIKF_KT, LIKF_NT2), *ws = xs + 256;
            auto stage = [&](int t) {
                const double *src = K.lik_ftab + (size_t)t * REC;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(wb + (size_t)(t & 1) * likf_rec_doubles(LIKF_KT, LIKF_NT2));
                for (int e = lane; e < REC / 2; e += 32)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 16u * e), "l"(src + 2 * e) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            double beta = 0., xe[NR];
            bool outside = false;
#pragma unroll
            for (int r = 0; r < NR; ++r) xe[r] = x_in[r];
            const double *hrec = K.lik_tab + (size_t)K.m * lik_rec_doubles(NR);
            if (EXT) lik_pre<NR>(msm, K.lx, K.n, lane, x_in, hrec, live, xe, outside, beta);
            __syncwarp();
            stage(0);
#pragma unroll
            for (int r = 0; r < NR; ++r) xs[gi_ * 32 + 4 * r + lg_] = xe[r];
            if (NR < 8) { for (int r = NR; r < 8; ++r) xs[gi_ * 32 + 4 * r + lg_] = 0.; }
            __syncwarp();
            double phi[LIKF_KT];
            likf_features(K.lik_fpt, K.kt1, lane, xs, phi);
            LikFeatAcc A;
            A.acc2 = 0.; A.S1 = 0.;
#pragma unroll
            for (int t2 = 0; t2 < LIKF_NT2; ++t2) A.w[t2][0] = A.w[t2][1] = 0.;
#pragma unroll 1
            for (int t = 0; t < K.nt1; ++t) {
                if (t + 1 < K.nt1) { stage(t + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                likf_tile(wb + (size_t)(t & 1) * likf_rec_doubles(LIKF_KT, LIKF_NT2), K.kt1, K.nt2, lane, phi, EXT && K.lx.use_bound, outside, beta,
                          K.lx.alpha, A);
                __syncwarp();
            }
            double acc2, S1;
            likf_finish<NR>(K.lik_fgt, K.nt2, K.n, lane, xs, ws, A, acc2, S1, gn);
            if (EXT) lik_post<NR>(msm, K.lx, K.n, lane, x_in, hrec, outside, beta, S1, acc2, K.e_c0, gn, lp);
            else lp = K.e_c0 - 0.5 * acc2;
            ke = qsum(ke_of(gn));
            return;
        }
        constexpr int NT4 = (NR + 1) / 2, REC = lik_rec_doubles(NR), OL = NR * NT4 * 32;
        const int lg_ = lane & 3;
        double *buf = const_cast<double *>(bsm) + (size_t)(threadIdx.x >> 5) * 2 * REC;
        auto stage = [&](int o) {
            const double *src = K.lik_tab + (size_t)o * REC;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(buf + (size_t)(o & 1) * REC);
#pragma unroll
            for (int i = 0; i < (REC / 2 + 31) / 32; ++i) {
                const int e = i * 32 + lane;
                if (e < REC / 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 16u * e), "l"(src + 2 * e) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        double acc2 = 0., S1 = 0., beta = 0.;
        bool outside = false;
        double xe[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) { gn[r] = 0.; xe[r] = x_in[r]; }
        const double *hrec = K.lik_tab + (size_t)K.m * REC;
        if (EXT) lik_pre<NR>(msm, K.lx, K.n, lane, x_in, hrec, live, xe, outside, beta);
        __syncwarp();
        stage(0);
#pragma unroll 1
        for (int o = 0; o < K.m; ++o) {
            if (o + 1 < K.m) { stage(o + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const double *rec = buf + (size_t)(o & 1) * REC;
            double a_[NT4][2];
#pragma unroll
            for (int t = 0; t < NT4; ++t) a_[t][0] = a_[t][1] = 0.;
#pragma unroll
            for (int kt = 0; kt < NR; ++kt)
#pragma unroll
                for (int t = 0; t < NT4; ++t) dmma884(a_[t][0], a_[t][1], xe[kt], rec[(kt * NT4 + t) * 32 + lane]);
            double fpart = 0., jr[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const double l_ = rec[OL + 4 * r + lg_], y_ = a_[r / 2][r % 2];
                fpart = fma(fma(0.5, y_, l_), xe[r], fpart);
                jr[r] = l_ + y_;
            }
            double f_ = rec[OL + 32] + qsum(fpart);
            if (EXT && K.lx.use_bound) {
                const double fmu = rec[OL + 33], f0 = f_;
                if (outside) f_ = (beta * f0 - (beta - K.lx.alpha) * fmu) / K.lx.alpha;
                S1 = fma(f_, (f0 - fmu) / K.lx.alpha, S1);
            }
            acc2 = fma(f_, f_, acc2);
#pragma unroll
            for (int r = 0; r < NR; ++r) gn[r] = fma(-f_, jr[r], gn[r]);
            __syncwarp();                                  // slot (o & 1) is refilled by the stage() of the next iteration
        }
        if (EXT) lik_post<NR>(msm, K.lx, K.n, lane, x_in, hrec, outside, beta, S1, acc2, K.e_c0, gn, lp);
        else lp = K.e_c0 - 0.5 * acc2;
        ke = qsum(ke_of(gn));
        return;
    }
    const double *mu_t = msm, *lin_t = msm + 32;
    const int lg = lane & 3;
    double acc[SH::NT][2];
#pragma unroll
    for (int t = 0; t < SH::NT; ++t) acc[t][0] = acc[t][1] = 0.;
    double d0[NR], x[NR], d2[NR], tj[NR], tjj[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int j = 4 * r + lg;
        double xo = x_in[r];
        tj[r] = 1.; tjj[r] = 0.; d2[r] = 0.;
        if (EXT) {
            if (K.use_transform && j < K.n) {
                // only hard-bounded coordinates need exp(): the others are affine (transforms/_constraint.pyx:133-215)
                const int hbj = (int)msm[224 + j];
                if (hbj != 0) to_original_1(x_in[r], msm[160 + j], msm[192 + j], hbj, xo, tj[r], tjj[r]);
                else { xo = msm[160 + j] + x_in[r] * msm[192 + j]; tj[r] = msm[192 + j]; }
            }
            d2[r] = (j < K.n) ? xo - msm[128 + j] : 0.;
            if (K.use_scales) xo = (xo - msm[64 + j]) / msm[96 + j];
            if (j >= K.n) xo = 0.;
        }
        x[r] = xo;
        d0[r] = (EXT && j >= K.n) ? 0. : xo - mu_t[j];
    }
    // ---- stage A: h = H d, beta^2 = d . h  (and the decay ellipsoid) ----
    dmma_tiles<NR, MV, 0, SH::O_X>(bsm, lane, d0, d2, d0, d0, acc);
    double bpart = 0., b2part = 0.;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        bpart = fma(d0[r], acc[r / 2][r % 2], bpart);
        if (EXT) b2part = fma(d2[r], acc[SH::O_D2 + r / 2][r % 2], b2part);
    }
    double beta2, beta2d = 0.;
    if (EXT) { double z0 = 0., z1 = 0.; qsum4(bpart, b2part, z0, z1, lane); beta2 = bpart; beta2d = b2part; }
    else beta2 = qsum(bpart);
    const bool outside = live && (beta2 > K.alpha2);
    // divisions by beta are multiplications by 1 / beta: the padded dimensions hold exact zeros and a zero numerator
    // sends the FP64 division to its slow path
    const double beta = sqrt(beta2), rbeta = 1. / beta;
    if (outside) {
#pragma unroll
        for (int r = 0; r < NR; ++r)
            x[r] = (4 * r + lg < K.n) ? (K.alpha * x[r] + (beta - K.alpha) * mu_t[4 * r + lg]) * rbeta : 0.;
    }
    // ---- stage B: the polynomial at x (inside) or at the projection x_0 (outside) ----
    double x2[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) x2[r] = x[r] * x[r];
    dmma_tiles<NR, MV, SH::O_X, SH::NT>(bsm, lane, x, x, x, x2, acc);
    double g3[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) g3[r] = 0.;
    if (SH::C3) {
        // ---- cubic-3: pair products of the (possibly projected) point times the pair-by-dimension coefficient matrix ----
        const double *t3 = msm + SH::MSM_DOUBLES;
        const int *pr = reinterpret_cast<const int *>(t3 + K.c3_kt * SH::N3T * 32);
        double *xs = reinterpret_cast<double *>(const_cast<int *>(pr) + K.c3_kt * 4) + (threadIdx.x >> 5) * 256 + (lane >> 2) * 32;
        __syncwarp();
#pragma unroll
        for (int r = 0; r < NR; ++r) xs[4 * r + lg] = x[r];
        __syncwarp();
        double a3[SH::N3T][2];
#pragma unroll
        for (int t = 0; t < SH::N3T; ++t) a3[t][0] = a3[t][1] = 0.;
        const double *b3 = t3 + lane;
#pragma unroll 2
        for (int kt = 0; kt < K.c3_kt; ++kt) {
            const int pk = pr[4 * kt + lg];
            const double a = xs[pk & 0xff] * xs[pk >> 8];
#pragma unroll
            for (int t = 0; t < SH::N3T; ++t) dmma884(a3[t][0], a3[t][1], a, b3[(kt * SH::N3T + t) * 32]);
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) g3[r] = a3[r / 2][r % 2];
    }
    double fpart = 0., jd = 0.;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const double y = acc[SH::O_X + r / 2][r % 2];
        const double lin_r = lin_t[4 * r + lg];
        double g = lin_r + y;
        fpart = fma(lin_r, x[r], fpart);
        fpart = fma(0.5 * x[r], y, fpart);
        if (C2) {
            const double t = acc[SH::O_X + (NR + r) / 2][(NR + r) % 2];
            const double u = acc[SH::O_X2 + r / 2][r % 2];
            g += fma(2. * x[r], t, u);
            fpart = fma(x2[r], t, fpart);
        }
        if (SH::C3) { g += g3[r]; fpart = fma(x[r] * (1. / 3.), g3[r], fpart); }     // Euler: sum_j x_j dP3/dx_j = 3 P3
        gn[r] = g;
        jd = fma(g, d0[r], jd);
    }
    double kp = EXT ? 0. : ke_of(gn), zz = 0.;
    qsum4(kp, jd, fpart, zz, lane);
    ke = kp;
    double fp = fpart;
    if (__any_sync(BFB_FULL, outside)) {
        // PolyModel._fj_bound, poly.py:480-503
        const double f0 = K.c0 + fpart;
        const double sfac = (f0 - K.f_mu) / K.alpha - jd * rbeta;
        double g2[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) g2[r] = outside ? gn[r] + sfac * (acc[r / 2][r % 2] * rbeta) : gn[r];
        double k2 = 0.;
        if (!EXT) k2 = qsum(ke_of(g2));
        if (outside) {
#pragma unroll
            for (int r = 0; r < NR; ++r) gn[r] = g2[r];
            fp = (beta * f0 - (beta - K.alpha) * K.f_mu) / K.alpha - K.c0;
            ke = k2;
        }
    }
    lp = K.c0 + fp;
    if (EXT) {
        // module rescale of the Jacobian, chain rule of the transform, decay, log-determinant of the transform
        double tpart = 0.;
        const bool dec = K.use_decay && (beta2d > K.d_alpha2);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int j = 4 * r + lg;
            double g = gn[r];
            if (K.use_scales) g = g / msm[96 + j];
            g *= tj[r];
            if (dec) g -= 2. * K.d_gamma * acc[SH::O_D2 + r / 2][r % 2];
            if (K.use_transform && j < K.n) {
                if ((int)msm[224 + j] != 0) { tpart += log(fabs(tj[r])); g += tjj[r] / tj[r]; }
                else tpart += msm[256 + j];            // tjj = 0: nothing to add to the gradient
            }
            gn[r] = (j < K.n) ? g : 0.;
        }
        if (K.use_decay) { const double ex = beta2d - K.d_alpha2; lp -= K.d_gamma * (ex > 0. ? ex : 0.); }
        double kq = ke_of(gn), z0 = 0., z1 = 0.;
        qsum4(kq, tpart, z0, z1, lane);
        ke = kq;
        if (K.use_transform) lp += tpart;
    }
}
