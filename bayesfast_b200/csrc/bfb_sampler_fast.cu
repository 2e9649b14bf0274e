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
#include "bfb_nuts_common.cuh"
#include <cstring>
#include <cstdlib>

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

// shared memory per chain, in doubles: XX (2 NP) | DD (NP) | TL q,p,g | TR q,p,g | PS | PB | stack levels (pl, pr, psum)
// | per-level scalars [5][10]: weight mantissa, weight exponent, proposal energy, proposal logp, proposal slot
__host__ __device__ inline int chain_smem_doubles(int NP, int LS) { return (11 + 3 * LS) * NP + 50; }

template <int G, int D, bool HAS_C2, int NK, int MB>
__global__ void __launch_bounds__(32, MB) nuts_multi_kernel(DevModel M, bfb_sampler_cfg cfg, ChainState st, RunOutDevF out,
                                                            int L, int LS, double *__restrict__ gstack,
                                                            double *__restrict__ gprop, int base_iter, int chunk_iters,
                                                            int n_groups, int n_units, int *__restrict__ queue)
{
    constexpr int CW = 32 / G;          // chains per warp
    constexpr int NP = G * D;           // padded dimension (16 or 32)
    constexpr int H2 = D / 2;           // 16-byte pairs per lane
    extern __shared__ double smem[];
    const int lane = threadIdx.x, gi = lane / G, lg = lane % G;
    // Persistent warps + ready ring.  A unit = (group of CW chains, chunk of `chunk_iters` iterations).  queue[0] = head,
    // queue[1] = tail, queue[2 .. 2 + n_groups) = chunks finished per group, then ring[n_units]: the ids of the groups
    // whose next chunk can start (initially every group once; a group is appended again when one of its chunks
    // completes).  Workers pop ready groups FIFO, so nobody waits for a particular predecessor and the SMs stay full
    // until the end of the launch instead of running 1.73 waves of whole-run blocks (4096 chains).
    volatile int *qv = queue;
    volatile int *ring = queue + 2 + n_groups;
#pragma unroll 1
    for (;;) {
    int idx = 0, group = 0;
    if (lane == 0) {
        idx = atomicAdd(queue, 1);
        if (idx < n_units) { while ((group = ring[idx]) < 0) __nanosleep(100); }
    }
    idx = __shfl_sync(BFB_FULL, idx, 0);
    if (idx >= n_units) break;
    group = __shfl_sync(BFB_FULL, group, 0);
    __threadfence();
    const int chunk = qv[2 + group];
    const int it_lo = chunk * chunk_iters;
    const int it_hi = min(out.n_iter, it_lo + chunk_iters);
    const int64_t c_raw = (int64_t)group * CW + gi;
    const bool exists = c_raw < st.C;
    const int64_t c = exists ? c_raw : st.C - 1;
    const int n = NK > 0 ? NK : M.n;
    double *csm = smem + (size_t)gi * chain_smem_doubles(NP, LS);
    double2 *XX = reinterpret_cast<double2 *>(csm);
    double *DD = csm + 2 * NP;
    const int oTL = 3 * NP, oTR = 6 * NP, oPS = 9 * NP, oPB = 10 * NP, oST = 11 * NP;
    double *ssc = csm + (11 + 3 * LS) * NP;
    double *gst = gstack + (size_t)c_raw * (size_t)(L > LS ? L - LS : 0) * 3 * NP;   // deep stack levels (L2 resident)
    // proposals (q, grad) are written once, at the leaf, into a slot of an L2-resident pool and are afterwards only
    // referred to by their slot index: merges move no vectors.  Same thread writes and reads a given element.
    double *gpr = gprop + (size_t)c * BFB_NSLOT * 2 * NP;

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

    const double c0 = M.c0[0];
    const double alpha = M.alpha, alpha2 = M.use_bound ? M.alpha * M.alpha : INFINITY;

    // ---- chain state (adaptation scalars stay in global memory: they are touched once per iteration) ----
    const size_t vb = (size_t)c * M.np;
    double q[D], p[D], g[D], var[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
        const int j = jdim(r);
        q[r] = st.q[vb + j]; g[r] = st.g[vb + j]; var[r] = st.var[vb + j]; p[r] = 0.;
    }
    RngF rng;
    rng.seed = cfg.seed; rng.chain = (uint64_t)(cfg.chain0 + c); rng.cached = ~0ull;
    int64_t t = st.t_draw[c];
    const int it0 = base_iter;                 // iterations done before this launch (same for every chain)
    double logp_q = st.logp[c];
    int status = exists ? st.status[c] : 9;
    unsigned tree_total = 0;
    bool done = (status != 0) || it_lo >= it_hi;
    int it = it_lo;

    // transition state
    double E0 = 0., step = 0., prop_E = 0., prop_lp = 0., acc_sum = 0., maxdE = 0.;
    WT Wtree; Wtree.m = 1.; Wtree.k = 0;
    int depth = 0, ileaf = 0, n_prop = 0, diverging = 0, prop_slot = 0, Rslot = 0;
    unsigned freemask = 0;
    double Rpl[D], Rps[D], REp = 0., Rlpp = 0.;
    WT RW; RW.m = 0.; RW.k = 0;
#pragma unroll
    for (int r = 0; r < D; ++r) Rpl[r] = Rps[r] = 0.;

    auto stack_ptr = [&](int lvl) -> double * {
        return (lvl < LS) ? (csm + oST + lvl * 3 * NP) : (gst + (size_t)(lvl - LS) * 3 * NP);
    };

    // ---- start of an iteration: base_hmc.py:62-80, Tree.__init__ nuts.py:27-43 ----
    auto start_iteration = [&](bool pred) {
        double p0[D], part = 0.;
#pragma unroll
        for (int r = 0; r < D; ++r) {
            const int j = jdim(r);
            p0[r] = (j < n) ? bfb_draw_normal(rng.seed, rng.chain, (uint64_t)(t + j)) / sqrt(var[r]) : 0.;
            part = fma(p0[r], var[r] * p0[r], part);
        }
        const double ke = gsum<G>(part);
        if (pred) {
            t += n;
            E0 = 0.5 * ke - logp_q;
            if (!isfinite(E0)) { status = 2; done = true; }
            const bool warm = (it0 + it) < cfg.n_warmup;
            const double eps = warm ? exp(st.log_step[c]) : exp(st.log_bar[c]);
            step = eps;
            VST(csm + oTL, q) VST(csm + oTL + NP, p0) VST(csm + oTL + 2 * NP, g)
            VST(csm + oTR, q) VST(csm + oTR + NP, p0) VST(csm + oTR + 2 * NP, g)
            VST(csm + oPS, p0)
            VST(gpr, q) VST(gpr + NP, g)                       // slot 0 = the starting point
            prop_slot = 0; freemask = ((1u << BFB_NSLOT) - 1u) & ~1u;
            prop_E = E0; prop_lp = logp_q; Wtree.m = 1.; Wtree.k = 0; acc_sum = 0.; maxdE = 0.;
            depth = 0; n_prop = 0; diverging = 0;
        }
    };
    // ---- start of a doubling: nuts.py:210 + the first lines of Tree.extend ----
    auto start_doubling = [&](bool pred) {
        const double ud = rng_uniform(rng, t);
        if (pred) {
            t += 1;
            const bool right = ud < 0.5;                     // log(u) < log(0.5) on the draw grid of bfb_rng.h
            const double *src = csm + (right ? oTR : oTL);
            VLD(q, src) VLD(p, src + NP) VLD(g, src + 2 * NP)
            VST(csm + oPB, p)
            step = right ? fabs(step) : -fabs(step);
            ileaf = 0;
        }
    };

    start_iteration(!done);
    start_doubling(!done);

#pragma unroll 1
    while (__any_sync(BFB_FULL, !done)) {
        const bool live = !done;
        // ================= leapfrog: integration.py:68-95 =================
        const double dt = 0.5 * step;
        double ph[D], x[D];
#pragma unroll
        for (int r = 0; r < D; ++r) {
            ph[r] = fma(dt, g[r], p[r]);
            x[r] = fma(step, var[r] * ph[r], q[r]);
        }
        double gn[D], fp = 0., ke2 = 0.;
        bool outside = false;
        double beta = 0., d0[D], hd0[D];
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            // pass 0: polynomial at the new point; pass 1 (only if some chain left the ellipsoid): at its projection
            __syncwarp();
#pragma unroll
            for (int r2 = 0; r2 < H2; ++r2) {
                const int j0 = r2 * 2 * G + 2 * lg;
                const double xa = x[2 * r2], xb = x[2 * r2 + 1];
                XX[j0] = make_double2(xa, xa * xa);
                XX[j0 + 1] = make_double2(xb, xb * xb);
                const double2 mu2 = M.use_bound ? __ldg(reinterpret_cast<const double2 *>(M.mu) + r2 * G + lg) : make_double2(0., 0.);
                reinterpret_cast<double2 *>(DD)[r2 * G + lg] = make_double2(xa - mu2.x, xb - mu2.y);
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
                const double lin = __ldg(M.lin + jdim(r));
                const double dr = x[r] - (M.use_bound ? __ldg(M.mu + jdim(r)) : 0.);
                gg[r] = lin + y[r];
                fpart = fma(lin, x[r], fpart);
                fpart = fma(0.5 * x[r], y[r], fpart);
                if (HAS_C2) { gg[r] += fma(2. * x[r], tt[r], u[r]); fpart = fma(x[r] * x[r], tt[r], fpart); }
                bpart = fma(dr, hh[r], bpart);
            }
            if (pass == 0) {
                double pn_part = 0.;
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    gn[r] = gg[r];
                    const double pn = fma(dt, gg[r], ph[r]);
                    pn_part = fma(pn, var[r] * pn, pn_part);
                }
                double zz = 0.;
                gsum4<G>(pn_part, bpart, fpart, zz, lane);
                fp = fpart; ke2 = pn_part;
                outside = live && (bpart > alpha2);
                if (!__any_sync(BFB_FULL, outside)) break;
                // PolyModel._fj_bound, poly.py:480-503: project onto the ellipsoid and evaluate there
                beta = sqrt(bpart);
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    const double mur = M.use_bound ? __ldg(M.mu + jdim(r)) : 0.;
                    d0[r] = x[r] - mur; hd0[r] = hh[r];
                    if (outside) x[r] = (jdim(r) < n) ? (alpha * x[r] + (beta - alpha) * mur) / beta : 0.;
                }
            } else {
                double jd_part = 0.;
#pragma unroll
                for (int r = 0; r < D; ++r) jd_part = fma(gg[r], d0[r], jd_part);
                double zz = 0., z2 = 0.;
                gsum4<G>(fpart, jd_part, zz, z2, lane);
                double pn_part = 0.;
                if (outside) {
                    const double f_mu = M.f_mu[0];
                    const double f0 = c0 + fpart;
                    const double sfac = (f0 - f_mu) / alpha - jd_part / beta;
#pragma unroll
                    for (int r = 0; r < D; ++r) {
                        gn[r] = gg[r] + sfac * (hd0[r] / beta);
                        const double pn = fma(dt, gn[r], ph[r]);
                        pn_part = fma(pn, var[r] * pn, pn_part);
                    }
                    fp = (beta * f0 - (beta - alpha) * f_mu) / alpha - c0;     // so that logp = c0 + fp below
                }
                pn_part = gsum<G>(pn_part);
                if (outside) ke2 = pn_part;
            }
        }
        const double lp = c0 + fp;
        const double E = 0.5 * ke2 - lp;
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
            for (int r = 0; r < D; ++r) { Rpl[r] = p[r]; Rps[r] = p[r]; }
            RW = wl; REp = E; Rlpp = lp;
            Rslot = __ffs(freemask) - 1;
            freemask &= ~(1u << Rslot);
            double *slot = gpr + (size_t)Rslot * 2 * NP;
            VST(slot, q) VST(slot + NP, g)
        }
        if (div_now) diverging = 1;
        // ================= merges: Tree._build_subtree, nuts.py:134-178 =================
        int lvl = 0;
        bool need = live && !div_now && ((ileaf >> lvl) & 1);
#pragma unroll 1
        while (__any_sync(BFB_FULL, need)) {
            const int lv = need ? lvl : 0;
            const double *sp = stack_ptr(lv);
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
                WT T1W; T1W.m = ssc[lv]; T1W.k = (int)ssc[10 + lv];
                const int T1slot = (int)ssc[40 + lv];
                const WT tot = wt_add(T1W, RW);
                if (!wt_select(um, tot, RW)) {               // keep tree1's proposal (nuts.py:164-167)
                    freemask |= 1u << Rslot;
                    Rslot = T1slot; REp = ssc[20 + lv]; Rlpp = ssc[30 + lv];
                } else {
                    freemask |= 1u << T1slot;
                }
#pragma unroll
                for (int r = 0; r < D; ++r) { Rpl[r] = T1pl[r]; Rps[r] = ps[r]; }
                RW = tot;
                if (turning) turn = true;
                lvl++;
            }
            need = need && !turn && ((ileaf >> lvl) & 1);
        }
        const bool fin = live && (div_now || turn || (ileaf + 1 == (1 << depth)));
        const bool push = live && !fin;
        if (__any_sync(BFB_FULL, push)) {
            double *sp = stack_ptr(push ? lvl : 0);
            if (push) {
                VST(sp, Rpl) VST(sp + NP, p) VST(sp + 2 * NP, Rps)
                ssc[lvl] = RW.m; ssc[10 + lvl] = (double)RW.k; ssc[20 + lvl] = REp; ssc[30 + lvl] = Rlpp;
                ssc[40 + lvl] = (double)Rslot;
                ileaf += 1;
            }
        }
        // ================= end of a doubling: Tree.extend, nuts.py:45-103 =================
        if (__any_sync(BFB_FULL, fin)) {
            const double ue = rng_uniform(rng, t);
            const bool right = step > 0.;
            if (fin) {
                double *dst = csm + (right ? oTR : oTL);
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
                if (right) {
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
                    freemask |= 1u << prop_slot;
                    prop_slot = Rslot; prop_E = REp; prop_lp = Rlpp;
                } else {
                    freemask |= 1u << Rslot;
                }
                Wtree = tot;
                VST(csm + oPS, PS)
                if (turning) turn = true;
            }
            const bool iter_end = fin && (div_now || turn || (depth >= cfg.max_treedepth));
            // ---------- end of the iteration: base_hmc.py:80-85 ----------
            if (__any_sync(BFB_FULL, iter_end)) {
                const bool warm = (it0 + it) < cfg.n_warmup;
                const double *slot = gpr + (size_t)(iter_end ? prop_slot : 0) * 2 * NP;
                double qn[D], gnx[D];
                VLD(qn, slot) VLD(gnx, slot + NP)
                const double accept_stat = acc_sum / (double)(n_prop > 0 ? n_prop : 1);
                // adaptation scalars live in global memory: EVERY lane of the chain reads them first, then (after a
                // warp barrier) lane 0 of the chain writes the updated values
                double log_step = st.log_step[c], log_bar = st.log_bar[c];
                const double hbar0 = st.hbar[c], mu_da = st.mu_da[c];
                const int64_t count = st.count[c];
                const int64_t n_samples = st.n_samples[c], previous_update = st.previous_update[c];
                const int adapt_window = st.adapt_window[c];
                const double fg_n = st.fg_n[c] + 1., bg_n = st.bg_n[c] + 1.;
                __syncwarp();
                if (iter_end) {
#pragma unroll
                    for (int r = 0; r < D; ++r) { q[r] = qn[r]; g[r] = gnx[r]; }
                    logp_q = prop_lp;
                    tree_total += (unsigned)n_prop;
                    if (warm && cfg.adapt_step_size) {      // step_size.py:31-45
                        const double cnt = (double)count;
                        const double w = 1. / (cnt + cfg.t0);
                        const double hbar = ((1. - w) * hbar0 + w * (cfg.target_accept - accept_stat));
                        log_step = mu_da - hbar * sqrt(cnt) / cfg.gamma;
                        const double mk = pow(cnt, -cfg.k);
                        log_bar = mk * log_step + (1. - mk) * log_bar;
                        if (lg == 0) { st.hbar[c] = hbar; st.log_step[c] = log_step; st.log_bar[c] = log_bar; st.count[c] = count + 1; }
                    }
                    if (warm && cfg.adapt_metric) {         // metrics.py:186-211, Welford state kept in global memory
                        const int64_t delta = n_samples - previous_update;
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
                            if (upd && j < n) var[r] = fgr / fg_n;
                            if (swap) { fgm = bgm; fgr = bgr; bgm = 0.; bgr = 0.; }
                            st.fg_mean[vb + j] = fgm; st.fg_raw[vb + j] = fgr;
                            st.bg_mean[vb + j] = bgm; st.bg_raw[vb + j] = bgr;
                        }
                        if (lg == 0) {
                            st.fg_n[c] = swap ? bg_n : fg_n;
                            st.bg_n[c] = swap ? 10. : bg_n;
                            if (swap) { st.previous_update[c] = n_samples; if (cfg.doubling) st.adapt_window[c] = adapt_window * 2; }
                            st.n_samples[c] = n_samples + 1;
                        }
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
                    if (it >= it_hi) done = true;
                }
                __syncwarp();      // the adaptation scalars written by lane 0 of a chain are re-read by all its lanes
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
            st.logp[c] = logp_q; st.t_draw[c] = t; st.iter[c] = it0 + it;
            st.status[c] = status;
            if (tree_total) atomicAdd(st.tree_total, (unsigned long long)tree_total);
        }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        qv[2 + group] = chunk + 1;
        if ((chunk + 1) * chunk_iters < out.n_iter) {
            const int ti = atomicAdd(queue + 1, 1);
            __threadfence();
            ring[ti] = group;
        }
    }
    }   // unit loop
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
    // L2-resident pools: deep stack levels (3 vectors each) followed by the proposal slots (q, grad per slot)
    const size_t deep = (size_t)(L > LS ? L - LS : 0) * 3 * NP;
    const size_t prop = (size_t)BFB_NSLOT * 2 * NP;
    if ((deep + prop) * (size_t)C > h->gstack_len) {
        if (h->gstack) cudaFree(h->gstack);
        h->gstack = nullptr; h->gstack_len = 0;
        BFB_CUDA(cudaMalloc((void **)&h->gstack, sizeof(double) * (deep + prop) * (size_t)C));
        h->gstack_len = (deep + prop) * (size_t)C;
    }
    RunOutDevF od;
    od.o = o; od.n_iter = n_iter;
    const int n_groups = (int)((C + CW - 1) / CW);
    int chunk_iters = (n_iter + 5) / 6;
    if (chunk_iters < 16) chunk_iters = n_iter < 16 ? n_iter : 16;
    if (const char *e = getenv("BFB200_CHUNK_ITERS")) { int v = atoi(e); if (v >= 1) chunk_iters = v; }
    const int n_chunks = (n_iter + chunk_iters - 1) / chunk_iters;
    const int64_t n_units64 = (int64_t)n_groups * n_chunks;
    BFB_REQUIRE(n_units64 < (1ll << 31), BFB_ERR_ARG, "too many work units");
    const size_t qlen = 2 + (size_t)n_groups + (size_t)n_units64;
    if (qlen > h->queue_len) {
        if (h->queue) cudaFree(h->queue);
        h->queue = nullptr; h->queue_len = 0;
        BFB_CUDA(cudaMalloc((void **)&h->queue, sizeof(int) * qlen));
        h->queue_len = qlen;
    }
    queue_init_kernel<<<(unsigned)((n_units64 + 255) / 256), 256, 0, h->stream>>>(h->queue, n_groups, (int)n_units64);
    h->launches++;
    int blocks = h->sm_count * MB;
    if (const char *e = getenv("BFB200_PERSISTENT")) { if (atoi(e) == 0) blocks = (int)n_units64; }
    if ((int64_t)blocks > n_units64) blocks = (int)n_units64;
    nuts_multi_kernel<G, D, HAS_C2, NK, MB><<<blocks, 32, smem, h->stream>>>(h->dm, h->scfg, h->cs, od, L, LS, h->gstack,
                                                                           h->gstack + deep * (size_t)C, (int)h->iters_done,
                                                                           chunk_iters, n_groups, (int)n_units64, h->queue);
    h->launches++;
    BFB_CUDA(cudaGetLastError());
    return BFB_OK;
}

// returns 1 if the fast path does not apply (caller uses the generic kernel), 0 on launch, <0 on error
int bfb_launch_nuts_fast(bfb_context *h, const bfb_run_out &o, int n_iter)
{
    const DevModel &M = h->dm;
    if (M.n > 32 || M.has_c3 || M.use_decay || M.use_transform || M.use_scales || M.epilogue) return 1;
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
