// bfb_sampler_fast.cu -- NUTS fast path: SEVERAL CHAINS PER WARP, asynchronous, predicated state machine.
//
// Applies to input_size <= 32, logp = output 0 of a linear + quadratic (+ cubic-2) PolyModel with the radial
// bound and no decay / variable transform / module rescale (the BASELINE.json headline configuration); anything
// else runs the generic warp-per-chain kernel of bfb_sampler.cu.  Same algorithm, same draw order.
//
// Why (profiles/r01_a, r01_b): with one chain per warp 78 % of the issue slots went to warp-uniform SCALAR work
// (tree bookkeeping, Philox, transcendentals, shuffle reductions) executed redundantly by 32 lanes, and the
// matrix-vector products had no ILP.  Here
//   * a chain is owned by G lanes (G = 8 or 16), each lane owning D = 32/G dimensions (n <= 16: 16/G); a warp
//     advances 32/G chains at once, so every scalar instruction serves 32/G chains and every lane carries D
//     independent FMA streams;
//   * the chains of a warp are NOT in lock step: each has its own iteration / tree depth / leaf counter and the
//     warp loops "one leapfrog for every live chain, then the post-processing each chain needs", every section
//     entered when any chain needs it and committed per chain by predication.  No chain ever waits for another,
//     so divergent tree depths cost nothing but the repeated merge section;
//   * coefficient rows come through L1 as 16-byte loads: the lanes of one chain read 128 contiguous bytes and the
//     other chains of the warp read the same addresses, i.e. one wavefront feeds 32/G chains (coefficients do not
//     fit in registers at this occupancy and per-lane shared-memory reads would cap the FP64 pipe at 25 %);
//   * multinomial weights are kept in the linear domain as (mantissa, exponent) pairs, so a merge needs no
//     exp / log / log1p: the reference's test log(u) < log w2 - log(w1 + w2) (nuts.py:82,164) becomes
//     u * (w1 + w2) < w2; one exp2 per leaf gives both the leaf weight and min(1, exp(-dE));
//   * U-turn dot products of a merge are reduced packed (8 values over the G lanes) and only signs are voted;
//   * Philox blocks are cached (two uniforms per block).
#include "bfb_common.cuh"
#include <cstring>

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

struct RngF {
    uint64_t seed, chain, cached;
    uint32_t w0, w1, w2, w3;
};
__device__ __forceinline__ double rng_uniform(RngF &r, int64_t t)
{
    const uint64_t blk = (uint64_t)t >> 1;
    if (blk != r.cached) {
        bfb_philox_block b = bfb_philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)r.chain,
                                               (uint32_t)(r.chain >> 32), (uint32_t)r.seed, (uint32_t)(r.seed >> 32));
        r.w0 = b.v[0]; r.w1 = b.v[1]; r.w2 = b.v[2]; r.w3 = b.v[3];
        r.cached = blk;
    }
    const uint64_t w = (t & 1) ? ((uint64_t)r.w2 | ((uint64_t)r.w3 << 32)) : ((uint64_t)r.w0 | ((uint64_t)r.w1 << 32));
    return bfb_u64_to_uniform(w);
}

__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(BFB_FULL, v, m); }

// sum over the G lanes of a chain (butterfly: bitwise identical on all its lanes)
template <int G>
__device__ __forceinline__ double gsum(double v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += shx(v, o);
    return v;
}

// four sums over the G lanes of a chain at once; every lane of the chain ends with all four totals
template <int G>
__device__ __forceinline__ void gsum4(double &a, double &b, double &c, double &d, int lane)
{
    const bool b0 = lane & 1, b1 = lane & 2;
    double k0 = b0 ? c : a, k1 = b0 ? d : b;
    double s0 = b0 ? a : c, s1 = b0 ? b : d;
    k0 += shx(s0, 1); k1 += shx(s1, 1);
    double k = b1 ? k1 : k0, s = b1 ? k0 : k1;
    k += shx(s, 2);
#pragma unroll
    for (int o = 4; o < G; o <<= 1) k += shx(k, o);
    const double o2 = shx(k, 2);
    const double p0 = b1 ? o2 : k, p1 = b1 ? k : o2;
    const double q0 = shx(p0, 1), q1 = shx(p1, 1);
    a = b0 ? q0 : p0; b = b0 ? q1 : p1; c = b0 ? p0 : q0; d = b0 ? p1 : q1;
}

// per chain: is any of the six sums over its G lanes <= 0 ?  (two padding slots hold +1)
template <int G>
__device__ __forceinline__ bool gany_nonpos6(double v0, double v1, double v2, double v3, double v4, double v5, int lane)
{
    const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4;
    double k0 = b0 ? v4 : v0, k1 = b0 ? v5 : v1, k2 = b0 ? 1. : v2, k3 = b0 ? 1. : v3;
    const double s0 = b0 ? v0 : v4, s1 = b0 ? v1 : v5, s2 = b0 ? v2 : 1., s3 = b0 ? v3 : 1.;
    k0 += shx(s0, 1); k1 += shx(s1, 1); k2 += shx(s2, 1); k3 += shx(s3, 1);
    double m0 = b1 ? k2 : k0, m1 = b1 ? k3 : k1;
    const double t0 = b1 ? k0 : k2, t1 = b1 ? k1 : k3;
    m0 += shx(t0, 2); m1 += shx(t1, 2);
    double k = b2 ? m1 : m0;
    const double s = b2 ? m0 : m1;
    k += shx(s, 4);
#pragma unroll
    for (int o = 8; o < G; o <<= 1) k += shx(k, o);
    const unsigned bal = __ballot_sync(BFB_FULL, k <= 0.);
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
    return (bal & gmask) != 0u;
}

// shared memory per chain, in doubles: XX (2 NP) | DD (NP) | TL q,p,g | TR q,p,g | PS | PQ | PG | PB | stack levels
// | per-level scalars [4][10]: weight mantissa, weight exponent, proposal energy, proposal logp
__host__ __device__ inline int chain_smem_doubles(int NP, int LS) { return (13 + 5 * LS) * NP + 40; }

template <int G, int D, bool HAS_C2, int NK, int MB>
__global__ void __launch_bounds__(32, MB) nuts_multi_kernel(DevModel M, bfb_sampler_cfg cfg, ChainState st, RunOutDevF out,
                                                        int L, int LS, double *__restrict__ gstack)
{
    constexpr int CW = 32 / G;          // chains per warp
    constexpr int NP = G * D;           // padded dimension (16 or 32)
    constexpr int H2 = D / 2;           // 16-byte pairs per lane
    extern __shared__ double smem[];
    const int lane = threadIdx.x, gi = lane / G, lg = lane % G;
    const int64_t c_raw = (int64_t)blockIdx.x * CW + gi;
    const bool exists = c_raw < st.C;
    const int64_t c = exists ? c_raw : st.C - 1;
    const int n = NK > 0 ? NK : M.n;
    double *csm = smem + (size_t)gi * chain_smem_doubles(NP, LS);
    double2 *XX = reinterpret_cast<double2 *>(csm);
    double *DD = csm + 2 * NP;
    const int oTL = 3 * NP, oTR = 6 * NP, oPS = 9 * NP, oPQ = 10 * NP, oPG = 11 * NP, oPB = 12 * NP, oST = 13 * NP;
    double *gst = gstack + (size_t)c_raw * (size_t)(L > LS ? L - LS : 0) * 5 * NP;   // deep stack levels (L2 resident)

    // ---- vector access: lane owns dims j = r2*2G + 2*lg + e ----
    auto jdim = [&](int r) { return (r >> 1) * 2 * G + 2 * lg + (r & 1); };
#define VLD(dst, base)                                                                      \
    _Pragma("unroll") for (int r2_ = 0; r2_ < H2; ++r2_) {                                  \
        const double2 v_ = reinterpret_cast<const double2 *>(base)[r2_ * G + lg];           \
        dst[2 * r2_] = v_.x; dst[2 * r2_ + 1] = v_.y;                                       \
    }
#define VST(base, src)                                                                      \
    _Pragma("unroll") for (int r2_ = 0; r2_ < H2; ++r2_)                                    \
        reinterpret_cast<double2 *>(base)[r2_ * G + lg] = make_double2(src[2 * r2_], src[2 * r2_ + 1]);
#define VSTP(pred, base, src) if (pred) { VST(base, src) }

    double lin[D], mu[D];
#pragma unroll
    for (int r = 0; r < D; ++r) { lin[r] = M.lin[jdim(r)]; mu[r] = M.use_bound ? M.mu[jdim(r)] : 0.; }
    const double c0 = M.c0[0];
    const double alpha = M.alpha, alpha2 = M.use_bound ? M.alpha * M.alpha : INFINITY;
    const double f_mu = M.use_bound ? M.f_mu[0] : 0.;

    // ---- chain state ----
    const size_t vb = (size_t)c * M.np;
    double q[D], p[D], g[D], var[D], inv_std[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
        const int j = jdim(r);
        q[r] = st.q[vb + j]; g[r] = st.g[vb + j]; var[r] = st.var[vb + j];
        inv_std[r] = 1. / sqrt(var[r]); p[r] = 0.;
    }
    RngF rng;
    rng.seed = cfg.seed; rng.chain = (uint64_t)(cfg.chain0 + c); rng.cached = ~0ull;
    int64_t t = st.t_draw[c];
    const int64_t it0 = st.iter[c];
    double logp_q = st.logp[c], fg_n = st.fg_n[c], bg_n = st.bg_n[c];
    double log_step = st.log_step[c], log_bar = st.log_bar[c], hbar = st.hbar[c];
    const double mu_da = st.mu_da[c];
    int64_t count = st.count[c], n_samples = st.n_samples[c], previous_update = st.previous_update[c];
    int adapt_window = st.adapt_window[c];
    int status = exists ? st.status[c] : 9;
    unsigned long long tree_total = 0;
    bool done = (status != 0) || out.n_iter <= 0;
    int it = 0;

    // transition state
    double E0 = 0., eps = 0., step = 0., prop_E = 0., prop_lp = 0., acc_sum = 0., maxdE = 0.;
    WT Wtree; Wtree.m = 1.; Wtree.k = 0;
    int depth = 0, dir = 1, ileaf = 0, nleaf = 1, n_prop = 0, diverging = 0;
    double Rpl[D], Rps[D], Rqp[D], Rgp[D], REp = 0., Rlpp = 0.;
    WT RW; RW.m = 0.; RW.k = 0;
#pragma unroll
    for (int r = 0; r < D; ++r) Rpl[r] = Rps[r] = Rqp[r] = Rgp[r] = 0.;

    // stack level address (shared for lvl < LS, global beyond)
    auto stack_ptr = [&](int lvl) -> double * {
        return (lvl < LS) ? (csm + oST + lvl * 5 * NP) : (gst + (size_t)(lvl - LS) * 5 * NP);
    };
    // per-level scalars of the pending left subtrees (L <= 10 on this path); every lane of a chain writes the
    // same value and later reads back what it wrote itself, so no synchronisation is needed
    double *ssc = csm + (13 + 5 * LS) * NP;

    // ---- start of an iteration: base_hmc.py:62-80, Tree.__init__ nuts.py:27-43 ----
    auto start_iteration = [&](bool pred) {
        double p0[D], part = 0.;
#pragma unroll
        for (int r = 0; r < D; ++r) {
            const int j = jdim(r);
            p0[r] = (j < n) ? inv_std[r] * bfb_draw_normal(rng.seed, rng.chain, (uint64_t)(t + j)) : 0.;
            part = fma(p0[r], var[r] * p0[r], part);
        }
        const double ke = gsum<G>(part);
        if (pred) {
            t += n;
            E0 = 0.5 * ke - logp_q;
            if (!isfinite(E0)) { status = 2; done = true; }
            const bool warm = (it0 + it) < cfg.n_warmup;
            eps = warm ? exp(log_step) : exp(log_bar);
            VST(csm + oTL, q) VST(csm + oTL + NP, p0) VST(csm + oTL + 2 * NP, g)
            VST(csm + oTR, q) VST(csm + oTR + NP, p0) VST(csm + oTR + 2 * NP, g)
            VST(csm + oPS, p0) VST(csm + oPQ, q) VST(csm + oPG, g)
            prop_E = E0; prop_lp = logp_q; Wtree.m = 1.; Wtree.k = 0; acc_sum = 0.; maxdE = 0.;
            depth = 0; n_prop = 0; diverging = 0;
        }
    };
    // ---- start of a doubling: nuts.py:210 + the first lines of Tree.extend ----
    auto start_doubling = [&](bool pred) {
        const double ud = rng_uniform(rng, t);
        if (pred) {
            t += 1;
            dir = (ud < 0.5) ? 1 : -1;       // log(u) < log(0.5) on the draw grid of bfb_rng.h
            const double *src = csm + (dir > 0 ? oTR : oTL);
            VLD(q, src) VLD(p, src + NP) VLD(g, src + 2 * NP)
            VST(csm + oPB, p)
            step = dir > 0 ? eps : -eps;
            ileaf = 0; nleaf = 1 << depth;
        }
    };

    start_iteration(!done);
    start_doubling(!done);

#pragma unroll 1
    while (__any_sync(BFB_FULL, !done)) {
        const bool live = !done;
        // ================= leapfrog: integration.py:68-95 =================
        const double dt = 0.5 * step;
        double ph[D], x[D], x2[D], dd[D];
#pragma unroll
        for (int r = 0; r < D; ++r) {
            ph[r] = fma(dt, g[r], p[r]);
            x[r] = fma(step, var[r] * ph[r], q[r]);
        }
        double gn[D], hd[D], fp = 0., beta2 = 0., ff0 = 0.;
        bool outside = false;
        double beta = 0., d0[D], hd0[D];
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            // pass 0: polynomial at the new point; pass 1 (only if some chain left the ellipsoid): at its projection
#pragma unroll
            for (int r = 0; r < D; ++r) { x2[r] = x[r] * x[r]; dd[r] = x[r] - mu[r]; }
            __syncwarp();
#pragma unroll
            for (int r2 = 0; r2 < H2; ++r2) {
                const int j0 = r2 * 2 * G + 2 * lg;
                XX[j0] = make_double2(x[2 * r2], x2[2 * r2]);
                XX[j0 + 1] = make_double2(x[2 * r2 + 1], x2[2 * r2 + 1]);
                reinterpret_cast<double2 *>(DD)[r2 * G + lg] = make_double2(dd[2 * r2], dd[2 * r2 + 1]);
            }
            __syncwarp();
            double y[D], tt[D], u[D], hh[D];
#pragma unroll
            for (int r = 0; r < D; ++r) { y[r] = 0.; tt[r] = 0.; u[r] = 0.; hh[r] = 0.; }
            const double2 *pS = reinterpret_cast<const double2 *>(M.S) + lg;
            const double2 *pH = reinterpret_cast<const double2 *>(M.HT) + lg;
            const double2 *pA1 = reinterpret_cast<const double2 *>(M.A1T) + lg;
            const double2 *pA2 = reinterpret_cast<const double2 *>(M.A2) + lg;
            // rows are 32 doubles = 16 double2 apart; with NK > 0 the loop is fully unrolled and every load has an
            // immediate offset from the four base pointers
#pragma unroll(NK > 0 ? NK : 2)
            for (int k = 0; k < n; ++k) {
                const double2 xk = XX[k];
                const double dk = DD[k];
#pragma unroll
                for (int r2 = 0; r2 < H2; ++r2) {
                    const int o2 = k * 16 + r2 * G;
                    const double2 s = __ldg(pS + o2);
                    const double2 h = __ldg(pH + o2);
                    y[2 * r2] = fma(s.x, xk.x, y[2 * r2]); y[2 * r2 + 1] = fma(s.y, xk.x, y[2 * r2 + 1]);
                    hh[2 * r2] = fma(h.x, dk, hh[2 * r2]); hh[2 * r2 + 1] = fma(h.y, dk, hh[2 * r2 + 1]);
                    if (HAS_C2) {
                        const double2 a1 = __ldg(pA1 + o2);
                        const double2 a2 = __ldg(pA2 + o2);
                        tt[2 * r2] = fma(a1.x, xk.x, tt[2 * r2]); tt[2 * r2 + 1] = fma(a1.y, xk.x, tt[2 * r2 + 1]);
                        u[2 * r2] = fma(a2.x, xk.y, u[2 * r2]); u[2 * r2 + 1] = fma(a2.y, xk.y, u[2 * r2 + 1]);
                    }
                }
            }
            double fpart = 0., bpart = 0.;
            double gg[D];
#pragma unroll
            for (int r = 0; r < D; ++r) {
                gg[r] = lin[r] + y[r];
                fpart = fma(lin[r], x[r], fpart);
                fpart = fma(0.5 * x[r], y[r], fpart);
                if (HAS_C2) { gg[r] += fma(2. * x[r], tt[r], u[r]); fpart = fma(x2[r], tt[r], fpart); }
                bpart = fma(dd[r], hh[r], bpart);
            }
            if (pass == 0) {
                double pn_part = 0.;
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    gn[r] = gg[r]; hd[r] = hh[r];
                    const double pn = fma(dt, gg[r], ph[r]);
                    pn_part = fma(pn, var[r] * pn, pn_part);
                }
                double zz = 0.;
                gsum4<G>(pn_part, bpart, fpart, zz, lane);
                fp = fpart; beta2 = bpart; ff0 = pn_part;      // ff0 temporarily holds 2*KE
                outside = live && (beta2 > alpha2);
                if (!__any_sync(BFB_FULL, outside)) break;
                // PolyModel._fj_bound, poly.py:480-503: project onto the ellipsoid and evaluate there
                beta = sqrt(beta2);
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    d0[r] = dd[r]; hd0[r] = hh[r];
                    if (outside) x[r] = (jdim(r) < n) ? (alpha * x[r] + (beta - alpha) * mu[r]) / beta : 0.;
                }
            } else {
                double jd_part = 0.;
#pragma unroll
                for (int r = 0; r < D; ++r) jd_part = fma(gg[r], d0[r], jd_part);
                double zz = 0., z2 = 0.;
                gsum4<G>(fpart, jd_part, zz, z2, lane);
                if (outside) {
                    const double f0 = c0 + fpart;
                    const double s = (f0 - f_mu) / alpha - jd_part / beta;
                    double pn_part = 0.;
#pragma unroll
                    for (int r = 0; r < D; ++r) {
                        gn[r] = gg[r] + s * (hd0[r] / beta);
                        const double pn = fma(dt, gn[r], ph[r]);
                        pn_part = fma(pn, var[r] * pn, pn_part);
                    }
                    fp = (beta * f0 - (beta - alpha) * f_mu) / alpha - c0;     // so that logp = c0 + fp below
                    pn_part = gsum<G>(pn_part);
                    ff0 = pn_part;
                } else {
                    (void)gsum<G>(0.);      // keep the warp's shuffles aligned
                }
            }
        }
        const double lp = c0 + fp;
        const double E = 0.5 * ff0 - lp;
        if (live) {
#pragma unroll
            for (int r = 0; r < D; ++r) {
                q[r] = fma(step, var[r] * ph[r], q[r]);
                g[r] = gn[r];
                p[r] = fma(dt, gn[r], ph[r]);
            }
        }
        // ================= leaf: Tree._single_step, nuts.py:105-132 =================
        double dE = E - E0;
        if (isnan(dE)) dE = INFINITY;
        bool div_now = false, turn = false;
        if (live) {
            if (fabs(dE) > fabs(maxdE)) maxdE = dE;
            n_prop += 1;
            div_now = !(fabs(dE) < cfg.max_change);
        }
        const WT wl = wt_from_dE(div_now ? 0. : dE);
        if (live && !div_now) {
            acc_sum += wt_min1(wl);
#pragma unroll
            for (int r = 0; r < D; ++r) { Rpl[r] = p[r]; Rps[r] = p[r]; Rqp[r] = q[r]; Rgp[r] = g[r]; }
            RW = wl; REp = E; Rlpp = lp;
        }
        if (div_now) diverging = 1;
        // ================= merges: Tree._build_subtree, nuts.py:134-178 =================
        int lvl = 0;
        bool need = live && !div_now && ((ileaf >> lvl) & 1);
#pragma unroll 1
        while (__any_sync(BFB_FULL, need)) {
            const double *sp = stack_ptr(need ? lvl : 0);
            double T1pl[D], T1pr[D], T1ps[D];
            VLD(T1pl, sp) VLD(T1pr, sp + NP) VLD(T1ps, sp + 2 * NP)
            double v0 = 0., v1 = 0., v2 = 0., v3 = 0., v4 = 0., v5 = 0.;
            double ps[D];
#pragma unroll
            for (int r = 0; r < D; ++r) {
                ps[r] = T1ps[r] + Rps[r];
                const double vT1pl = var[r] * T1pl[r], vp = var[r] * p[r];
                const double ps1 = T1ps[r] + Rpl[r], ps2 = T1pr[r] + Rps[r];
                v0 = fma(ps[r], vT1pl, v0); v1 = fma(ps[r], vp, v1);
                v2 = fma(ps1, vT1pl, v2); v3 = fma(ps1, var[r] * Rpl[r], v3);
                v4 = fma(ps2, var[r] * T1pr[r], v4); v5 = fma(ps2, vp, v5);
            }
            if (lvl < 1) { v2 = v3 = v4 = v5 = 1.; }         // extra checks only when depth > 1 (nuts.py:154)
            const bool turning = gany_nonpos6<G>(v0, v1, v2, v3, v4, v5, lane);
            const double um = rng_uniform(rng, t);
            if (need) {
                t += 1;
                WT T1W; T1W.m = ssc[lvl]; T1W.k = (int)ssc[10 + lvl];
                const WT tot = wt_add(T1W, RW);
                if (!wt_select(um, tot, RW)) {               // keep tree1's proposal (nuts.py:164-167)
                    VLD(Rqp, sp + 3 * NP) VLD(Rgp, sp + 4 * NP)
                    REp = ssc[20 + lvl]; Rlpp = ssc[30 + lvl];
                }
#pragma unroll
                for (int r = 0; r < D; ++r) { Rpl[r] = T1pl[r]; Rps[r] = ps[r]; }
                RW = tot;
                if (turning) turn = true;
                lvl++;
            }
            need = need && !turn && ((ileaf >> lvl) & 1);
        }
        const bool fin = live && (div_now || turn || (ileaf + 1 == nleaf));
        const bool push = live && !fin;
        if (__any_sync(BFB_FULL, push)) {
            double *sp = stack_ptr(push ? lvl : 0);
            if (push) {
                VST(sp, Rpl) VST(sp + NP, p) VST(sp + 2 * NP, Rps) VST(sp + 3 * NP, Rqp) VST(sp + 4 * NP, Rgp)
                ssc[lvl] = RW.m; ssc[10 + lvl] = (double)RW.k; ssc[20 + lvl] = REp; ssc[30 + lvl] = Rlpp;
                ileaf += 1;
            }
        }
        // ================= end of a doubling: Tree.extend, nuts.py:45-103 =================
        if (__any_sync(BFB_FULL, fin)) {
            bool stop = false;
            const double ue = rng_uniform(rng, t);
            if (fin) {
                double *dst = csm + (dir > 0 ? oTR : oTL);
                VST(dst, q) VST(dst + NP, p) VST(dst + 2 * NP, g)
                depth += 1;
            }
            const bool ok = fin && !div_now && !turn;
            double PS[D], PB[D], TLp[D], TRp[D];
            VLD(PS, csm + oPS) VLD(PB, csm + oPB) VLD(TLp, csm + oTL + NP) VLD(TRp, csm + oTR + NP)
            double v0 = 0., v1 = 0., v2 = 0., v3 = 0., v4 = 0., v5 = 0.;
#pragma unroll
            for (int r = 0; r < D; ++r) {
                PS[r] += Rps[r];
                const double vp = var[r] * p[r], vRpl = var[r] * Rpl[r], vPB = var[r] * PB[r];
                const double vTL = var[r] * TLp[r], vTR = var[r] * TRp[r];
                v0 = fma(PS[r], vTL, v0); v1 = fma(PS[r], vTR, v1);
                // nuts.py:86-98: self.p_sum is updated in place BEFORE p_sum1 / p_sum2 are formed, so the
                // "old tree" p_sum entering them is already the total (see the oracle, bf_oracle.c tree_extend)
                if (dir > 0) {
                    const double ps1 = PS[r] + Rpl[r], ps2 = PB[r] + Rps[r];
                    v2 = fma(ps1, vTL, v2); v3 = fma(ps1, vRpl, v3); v4 = fma(ps2, vPB, v4); v5 = fma(ps2, vp, v5);
                } else {
                    const double ps1 = Rps[r] + PB[r], ps2 = Rpl[r] + PS[r];
                    v2 = fma(ps1, vp, v2); v3 = fma(ps1, vPB, v3); v4 = fma(ps2, vRpl, v4); v5 = fma(ps2, vTR, v5);
                }
            }
            const bool turning = gany_nonpos6<G>(v0, v1, v2, v3, v4, v5, lane);
            if (ok) {
                t += 1;
                const WT tot = wt_add(Wtree, RW);
                if (wt_select(ue, Wtree, RW)) {               // nuts.py:81-83 biased progressive: log(u) < size2 - size1
                    VST(csm + oPQ, Rqp) VST(csm + oPG, Rgp)
                    prop_E = REp; prop_lp = Rlpp;
                }
                Wtree = tot;
                VST(csm + oPS, PS)
                if (turning) turn = true;
            }
            if (fin) stop = div_now || turn || (depth >= cfg.max_treedepth);
            const bool iter_end = fin && stop;
            // ---------- end of the iteration: base_hmc.py:80-85 ----------
            if (__any_sync(BFB_FULL, iter_end)) {
                const bool warm = (it0 + it) < cfg.n_warmup;
                double qn[D], gnx[D];
                VLD(qn, csm + oPQ) VLD(gnx, csm + oPG)
                const double accept_stat = acc_sum / (double)(n_prop > 0 ? n_prop : 1);
                if (iter_end) {
#pragma unroll
                    for (int r = 0; r < D; ++r) { q[r] = qn[r]; g[r] = gnx[r]; }
                    logp_q = prop_lp;
                    tree_total += (unsigned long long)n_prop;
                    if (warm && cfg.adapt_step_size) {      // step_size.py:31-45
                        const double cnt = (double)count;
                        const double w = 1. / (cnt + cfg.t0);
                        hbar = ((1. - w) * hbar + w * (cfg.target_accept - accept_stat));
                        log_step = mu_da - hbar * sqrt(cnt) / cfg.gamma;
                        const double mk = pow(cnt, -cfg.k);
                        log_bar = mk * log_step + (1. - mk) * log_bar;
                        count += 1;
                    }
                    if (warm && cfg.adapt_metric) {         // metrics.py:186-211, Welford state kept in global memory
                        const int64_t delta = n_samples - previous_update;
                        fg_n += 1.; bg_n += 1.;
                        const bool upd = ((delta + 1) % cfg.update_window == 0);
                        const bool swap = delta >= adapt_window;
#pragma unroll
                        for (int r = 0; r < D; ++r) {
                            const int j = jdim(r);
                            double fgm = st.fg_mean[vb + j], fgr = st.fg_raw[vb + j];
                            double bgm = st.bg_mean[vb + j], bgr = st.bg_raw[vb + j];
                            double od = q[r] - fgm;
                            fgm += od / fg_n;
                            fgr += 1. * od * (q[r] - fgm);
                            od = q[r] - bgm;
                            bgm += od / bg_n;
                            bgr += 1. * od * (q[r] - bgm);
                            if (upd && j < n) { var[r] = fgr / fg_n; inv_std[r] = 1. / sqrt(var[r]); }
                            if (swap) { fgm = bgm; fgr = bgr; bgm = 0.; bgr = 0.; }
                            st.fg_mean[vb + j] = fgm; st.fg_raw[vb + j] = fgr;
                            st.bg_mean[vb + j] = bgm; st.bg_raw[vb + j] = bgr;
                        }
                        if (swap) {
                            fg_n = bg_n; bg_n = 10.;
                            previous_update = n_samples;
                            if (cfg.doubling) adapt_window *= 2;
                        }
                        n_samples += 1;
                    }
                    const size_t o = (size_t)c * out.n_iter + it;
                    if (out.o.samples) {
#pragma unroll
                        for (int r = 0; r < D; ++r) if (jdim(r) < n) out.o.samples[o * n + jdim(r)] = q[r];
                    }
                    if (lg == 0) {
                        if (out.o.logp) out.o.logp[o] = prop_lp;
                        if (out.o.energy) out.o.energy[o] = prop_E;
                        if (out.o.tree_depth) out.o.tree_depth[o] = depth;
                        if (out.o.tree_size) out.o.tree_size[o] = n_prop;
                        if (out.o.mean_tree_accept) out.o.mean_tree_accept[o] = accept_stat;
                        if (out.o.step_size) out.o.step_size[o] = exp(log_step);
                        if (out.o.step_size_bar) out.o.step_size_bar[o] = exp(log_bar);
                        if (out.o.energy_change) out.o.energy_change[o] = prop_E - E0;
                        if (out.o.max_energy_change) out.o.max_energy_change[o] = maxdE;
                        if (out.o.diverging) out.o.diverging[o] = diverging;
                    }
                    it += 1;
                    if (it >= out.n_iter) done = true;
                }
                start_iteration(iter_end && !done);
            }
            start_doubling(fin && !done);
        }
    }

    // ---- persist chain state ----
    if (exists && st.status[c] == 0) {
#pragma unroll
        for (int r = 0; r < D; ++r) {
            const int j = jdim(r);
            st.q[vb + j] = q[r]; st.g[vb + j] = g[r]; st.var[vb + j] = var[r];
        }
        if (lg == 0) {
            st.logp[c] = logp_q; st.fg_n[c] = fg_n; st.bg_n[c] = bg_n;
            st.log_step[c] = log_step; st.log_bar[c] = log_bar; st.hbar[c] = hbar;
            st.count[c] = count; st.n_samples[c] = n_samples; st.previous_update[c] = previous_update;
            st.adapt_window[c] = adapt_window; st.t_draw[c] = t; st.iter[c] = it0 + it;
            st.status[c] = status;
            if (tree_total) atomicAdd(st.tree_total, tree_total);
        }
    }
}

template <int G, int D, bool HAS_C2, int NK, int MB = 8>
static int launch_multi(bfb_context *h, const bfb_run_out &o, int n_iter)
{
    constexpr int CW = 32 / G, NP = G * D;
    const int L = h->scfg.max_treedepth;
    // Shared memory vs L1: the coefficient tables (4 n x 32 doubles, 27 KB at n = 26) are read through L1, which
    // shares its 228 KB with shared memory, so the per-chain shared state is kept to ~100 KB per SM (MB resident
    // one-warp blocks): the first LS levels of the tree stack live in shared memory, deeper (rarely touched) levels
    // in an L2-resident global buffer.  (Padding shared memory to force 7 blocks per SM -- two equal waves for 2048
    // blocks -- was measured slower: it squeezes the tables out of L1.)
    const size_t budget = (size_t)(100 * 1024) / MB;
    int LS = L;
    while (LS > 1 && sizeof(double) * CW * chain_smem_doubles(NP, LS) > budget) --LS;
    if (const char *e = getenv("BFB200_STACK_LEVELS_SMEM")) { int v = atoi(e); if (v >= 1 && v <= L) LS = v; }
    const size_t smem = sizeof(double) * CW * chain_smem_doubles(NP, LS);
    BFB_CUDA(cudaFuncSetAttribute(nuts_multi_kernel<G, D, HAS_C2, NK, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t C = h->cs.C;
    const size_t deep = (size_t)(L > LS ? L - LS : 0) * 5 * NP;
    if (deep * (size_t)C > h->gstack_len) {
        if (h->gstack) cudaFree(h->gstack);
        h->gstack = nullptr; h->gstack_len = 0;
        BFB_CUDA(cudaMalloc((void **)&h->gstack, sizeof(double) * deep * (size_t)C));
        h->gstack_len = deep * (size_t)C;
    }
    RunOutDevF od;
    od.o = o; od.n_iter = n_iter;
    const int blocks = (int)((C + CW - 1) / CW);
    nuts_multi_kernel<G, D, HAS_C2, NK, MB><<<blocks, 32, smem, h->stream>>>(h->dm, h->scfg, h->cs, od, L, LS, h->gstack);
    h->launches++;
    BFB_CUDA(cudaGetLastError());
    return BFB_OK;
}

// returns 1 if the fast path does not apply (caller uses the generic kernel), 0 on launch, <0 on error
int bfb_launch_nuts_fast(bfb_context *h, const bfb_run_out &o, int n_iter)
{
    const DevModel &M = h->dm;
    if (M.n > 32 || M.has_c3 || M.use_decay || M.use_transform || M.use_scales) return 1;
    if (h->scfg.max_treedepth > 10) return 1;
    const int n = M.n;
    const char *gsel = getenv("BFB200_LANES_PER_CHAIN");
    const int G = gsel ? atoi(gsel) : 0;
    if (n <= 16) {
        // 8 lanes x 2 dims: four chains per warp
        if (M.has_c2) return n == 16 ? launch_multi<8, 2, true, 16>(h, o, n_iter) : launch_multi<8, 2, true, 0>(h, o, n_iter);
        return launch_multi<8, 2, false, 0>(h, o, n_iter);
    }
    if (G == 8) return M.has_c2 ? launch_multi<8, 4, true, 0>(h, o, n_iter) : launch_multi<8, 4, false, 0>(h, o, n_iter);
    // 16 lanes x 2 dims: two chains per warp
    if (M.has_c2 && n == 26) {
        const char *mb = getenv("BFB200_MINBLOCKS");
        const int MBv = mb ? atoi(mb) : 8;
        if (MBv == 12) return launch_multi<16, 2, true, 26, 12>(h, o, n_iter);
        return launch_multi<16, 2, true, 26, 8>(h, o, n_iter);
    }
    if (M.has_c2) return launch_multi<16, 2, true, 0>(h, o, n_iter);
    return launch_multi<16, 2, false, 0>(h, o, n_iter);
}
