// bfb_team.cuh -- PolyModel value + gradient for EIGHT chains evaluated by a TEAM of four warps (FP64 tensor cores).
//
// Why a team: with one warp per 8-chain group (bfb_dmma.cuh) 4096 chains are 512 warps on the 592 schedulers of a B200 --
// less than one warp per scheduler, 255 registers each, every dependent instruction exposes its latency (profiles/r01_s:
// issue slots 16 % busy, DMMA sub-pipe 16 %).  Here the four matrix-vector products of the cubic-2 stack are split over the
// OUTPUT dimensions: warp w of the team owns the dimensions j = 4 r + lg with r in [NRW w, NRW (w + 1)) (lg = lane & 3,
// the 8 chains are still the 8 rows of the m8n8k4 DMMA, a chain is still the quad lane >> 2) and computes only the output
// tiles of those dimensions -- a quarter of the DMMAs, a quarter of every state vector (2 registers instead of 7 at n = 26),
// ~1/2 of the registers, four times the warps.  What crosses the team goes through shared memory and named barriers
// (bar.sync id, 128):
//   * the point: every warp writes its slice of q into the team's x buffer (row r at r * 32 + lane) and reads all of it
//     back as the A operand (element kt * 32 + lane IS the A fragment of k-tile kt);
//   * the scalar reductions (Mahalanobis radius; value, J.d and kinetic energy): reduce-scatter over the quad with three
//     shuffles, one conflict-free store per lane, barrier, the four partials summed in a fixed order by everybody
//     (bitwise identical totals in all four warps, deterministic).
// Three barriers per evaluation.  The kinetic energy of the momentum after the second half kick is reduced together with
// the value; outside the radial bound (poly.py:480-503) the gradient gets the term sfac * H(x - mu) / beta, whose
// contribution to the kinetic energy is added algebraically (two more sums that ride along in the packed reductions)
// instead of through a fourth barrier.
//
// Operand table (host-built, bfb_upload_model): tfrag[w][kt][tile][lane], tiles of warp w ordered D | x block | x^2 block
// with the value index v = 2 tile + e of column 2 lg + e (v < NRW: first matrix of the block, dimension r = NRW w + v;
// NRW <= v < 2 NRW: second matrix of the x block).
#pragma once
#include "bfb_dmma.cuh"

#ifndef BFB_TEAM_C3_NACC
#define BFB_TEAM_C3_NACC 1       // accumulator sets of the cubic-3 stage (team_logp_grad); measured with 2: evaluation kernel +4 %, NUTS -14 %
#endif

template <int NR, int MV>
struct TeamShape {
    static constexpr bool C2 = MV & 1;
    static constexpr bool C3 = (MV & 4) != 0;                // cubic-3 configs: the pair-product GEMM, operand streamed from L2
    static constexpr int T = 4;                              // warps of a team
    static constexpr int NRW = (NR + 3) / 4;                 // dimensions per lane and warp
    static constexpr int NRP = 4 * NRW;                      // rows of a vector slot (>= NR; the padding rows hold zeros)
    static constexpr int SLOT = NRP * 32;
    static constexpr int TD = (NRW + 1) / 2;                 // tiles multiplying x - mu
    static constexpr int TX = C2 ? NRW : (NRW + 1) / 2;      // tiles multiplying x
    static constexpr int T2 = C2 ? (NRW + 1) / 2 : 0;        // tiles multiplying x^2
    static constexpr int NTW = TD + TX + T2;                 // tiles per warp and k-tile
    static constexpr int TAB_DOUBLES = T * NR * NTW * 32;
    static constexpr int TN3 = C3 ? (NRW + 1) / 2 : 0;       // output tiles of the cubic-3 GEMM per warp
    static constexpr int MSM_HALF = 4 * NRP > 32 ? 4 * NRP : 32;   // per-dimension tables staged behind the operand table: mu | lin
    static constexpr int MSM = 2 * MSM_HALF;
    static constexpr int RED_DOUBLES = 2 * 128;              // two alternating reduction buffers [w][chain][4]
};
inline int bfb_team_tiles(int nr, bool c2) { const int nrw = (nr + 3) / 4; return (nrw + 1) / 2 + (c2 ? nrw + (nrw + 1) / 2 : (nrw + 1) / 2); }

__device__ __forceinline__ void team_bar(int id) { asm volatile("bar.sync %0, 128;" :: "r"(id) : "memory"); }

// reduce-scatter of four sums over the 4 lanes of a quad: lane lg ends with the total of quantity ((lg & 1) << 1) | (lg >> 1)
__device__ __forceinline__ double qrs4(double a, double b, double c, double d, int lane)
{
    const bool b0 = lane & 1, b1 = lane & 2;
    double k0 = b0 ? c : a, k1 = b0 ? d : b;
    const double s0 = b0 ? a : c, s1 = b0 ? b : d;
    k0 += shx4(s0, 1); k1 += shx4(s1, 1);
    double k = b1 ? k1 : k0;
    const double s = b1 ? k0 : k1;
    k += shx4(s, 2);
    return k;
}

// per chain: is any of the six sums over its 4 lanes <= 0 ?  (U-turn tests, packed reduction over the quad)
__device__ __forceinline__ bool team_any_nonpos6(double v0, double v1, double v2, double v3, double v4, double v5, int lane)
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

// Four sums over all dimensions of each of the 8 chains of the team.  red: the team's RED_DOUBLES, rbuf: which half is
// written next (flips).  Every lane of every warp returns the same totals a, b; c and d are summed by the leader warp
// (w == 0) only -- FP64 additions cost pipe cycles that the DMMAs need (the other warps get their own partials back).
__device__ __forceinline__ void team_sum4(double &a, double &b, double &c, double &d, double *red, int &rbuf, int bar_id, int lane, int w)
{
    const int lg = lane & 3, gi = lane >> 2;
    const double k = qrs4(a, b, c, d, lane);
    double *buf = red + rbuf * 128;
    buf[w * 32 + gi * 4 + (((lg & 1) << 1) | (lg >> 1))] = k;
    team_bar(bar_id);
    const double2 *rp = reinterpret_cast<const double2 *>(buf + gi * 4);
    const double2 u0 = rp[0], u1 = rp[16], u2 = rp[32], u3 = rp[48];
    a = (u0.x + u1.x) + (u2.x + u3.x);
    b = (u0.y + u1.y) + (u2.y + u3.y);
    if (w == 0) {
        const double2 v0 = rp[1], v1 = rp[17], v2 = rp[33], v3 = rp[49];
        c = (v0.x + v1.x) + (v2.x + v3.x);
        d = (v0.y + v1.y) + (v2.y + v3.y);
    }
    rbuf ^= 1;
}

// Evaluation at the point whose slices the owners have written to the team's exchange buffers: xb[0 .. SLOT) = x,
// xb[SLOT .. 2 SLOT) = x - mu, xb[2 SLOT .. 3 SLOT) = x^2 (row r at r * 32 + lane; the caller has passed a team barrier
// since).  tab_w: this warp's slice of the operand table; q_own / p / var: own dimensions (p = momentum after the first half
// kick).  Returns the gradient of the own dimensions in every warp; lp and ke = sum_j var_j (p_j + dt g_j)^2 over ALL
// dimensions of the chain are valid in the leader warp (w == 0) only.
// Outside the radial bound the owners replace their slices of x and x^2 by the projection x_0 = mu + (alpha / beta)(x - mu)
// (poly.py:480-503 writes (alpha x + (beta - alpha) mu) / beta) behind one more barrier, so that nobody recomputes the whole
// point; rounds in which no chain of the team is outside skip it.
template <int NR, int MV>
__device__ __forceinline__ void team_logp_grad(const double *tab_w, const double *msm, double *xb, double *red, int &rbuf,
                                               int bar_id, int lane, int w, const DmmaConsts &K, bool live,
                                               const double (&q_own)[TeamShape<NR, MV>::NRW], const double (&p)[TeamShape<NR, MV>::NRW],
                                               const double (&var)[TeamShape<NR, MV>::NRW], double dt,
                                               double &lp, double (&gn)[TeamShape<NR, MV>::NRW], double &ke)
{
    using TS = TeamShape<NR, MV>;
    constexpr int NRW = TS::NRW, NTW = TS::NTW, TD = TS::TD, TX = TS::TX, T2 = TS::T2, SLOT = TS::SLOT;
    constexpr bool C2 = TS::C2;
    const int lg = lane & 3;
    const double *mu_t = msm, *lin_t = msm + TS::MSM_HALF;
    const double *tb = tab_w + lane;
    double mu_o[NRW], d_o[NRW];
#pragma unroll
    for (int i = 0; i < NRW; ++i) { mu_o[i] = mu_t[4 * (NRW * w + i) + lg]; d_o[i] = q_own[i] - mu_o[i]; }
    // ---- stage A: h = H (x - mu) for the own dimensions; two accumulator sets (even / odd k-tiles) halve the dependent chain ----
    double ah[TD][2][2];
#pragma unroll
    for (int t = 0; t < TD; ++t) ah[t][0][0] = ah[t][0][1] = ah[t][1][0] = ah[t][1][1] = 0.;
#pragma unroll
    for (int kt = 0; kt < NR; ++kt) {
        const double dk = xb[SLOT + kt * 32 + lane];
#pragma unroll
        for (int t = 0; t < TD; ++t) dmma884(ah[t][kt & 1][0], ah[t][kt & 1][1], dk, tb[(kt * NTW + t) * 32]);
    }
    double h[NRW], bpart = 0., vh2 = 0., z0 = 0., z1 = 0.;
#pragma unroll
    for (int i = 0; i < NRW; ++i) {
        h[i] = ah[i / 2][0][i % 2] + ah[i / 2][1][i % 2];
        bpart = fma(d_o[i], h[i], bpart);
        vh2 = fma(var[i] * h[i], h[i], vh2);
    }
    team_sum4(bpart, z0, vh2, z1, red, rbuf, bar_id, lane, w);
    const double beta2 = bpart;
    const bool outside = live && (beta2 > K.alpha2);
    double rbeta = 0., beta = 0.;
    double xo[NRW];
#pragma unroll
    for (int i = 0; i < NRW; ++i) xo[i] = q_own[i];
    if (__any_sync(BFB_FULL, outside)) {           // identical in the four warps: same totals, same flags
        rbeta = rsqrt(beta2); beta = beta2 * rbeta;
        const double sc = K.alpha * rbeta;
#pragma unroll
        for (int i = 0; i < NRW; ++i) {
            if (outside) xo[i] = (4 * (NRW * w + i) + lg < K.n) ? fma(sc, d_o[i], mu_o[i]) : 0.;
            xb[(NRW * w + i) * 32 + lane] = xo[i];
            xb[2 * SLOT + (NRW * w + i) * 32 + lane] = xo[i] * xo[i];
        }
        team_bar(bar_id);
    }
    // ---- stage B: the polynomial at x (inside) or at its projection (outside) ----
    double ax[TX][2], a2[T2 > 0 ? T2 : 1][2];
#pragma unroll
    for (int t = 0; t < TX; ++t) ax[t][0] = ax[t][1] = 0.;
#pragma unroll
    for (int t = 0; t < (T2 > 0 ? T2 : 1); ++t) a2[t][0] = a2[t][1] = 0.;
#pragma unroll
    for (int kt = 0; kt < NR; ++kt) {
        const double xk = xb[kt * 32 + lane];
#pragma unroll
        for (int t = 0; t < TX; ++t) dmma884(ax[t][0], ax[t][1], xk, tb[(kt * NTW + TD + t) * 32]);
        if (C2) {
            const double x2k = xb[2 * SLOT + kt * 32 + lane];
#pragma unroll
            for (int t = 0; t < T2; ++t) dmma884(a2[t][0], a2[t][1], x2k, tb[(kt * NTW + TD + TX + t) * 32]);
        }
    }
    // ---- cubic-3 (MV bit 2): gradient of P3 = sum_(j<k<l) a_jkl x_j x_k x_l as the GEMM [8 chains x pairs (k < l)] . [pairs x n]
    // of the pair products of the (possibly projected) point, like bfb_dmma.cuh -- but at n = 64 the operand is 1 MB (2016
    // pairs x 64 dimensions), so it streams from L2: every warp reads only the columns of its own dimensions (TN3 tiles per
    // k-tile, table layout [kt][warp][tile][lane]), U k-tiles of operand and pair indices are loaded while the previous U run
    // on the tensor cores.  Same accumulation order as the shared-memory version (k-tiles in sequence).
    double g3[NRW];
#pragma unroll
    for (int i = 0; i < NRW; ++i) g3[i] = 0.;
    if constexpr (TS::C3) {
        constexpr int TN3 = TS::TN3, U = 8, NACC = BFB_TEAM_C3_NACC;
        double a3[NACC][TN3][2];                                // k-tile kt accumulates into set kt mod NACC: shorter dependent DMMA chains
#pragma unroll
        for (int s_ = 0; s_ < NACC; ++s_)
#pragma unroll
            for (int t = 0; t < TN3; ++t) a3[s_][t][0] = a3[s_][t][1] = 0.;
        const double *xq = xb + (lane & ~3);                  // the quad of this lane's chain inside a row of the exchange buffer
        const double *t3 = K.team3 + (size_t)w * TN3 * 32 + lane;
        const int *pr = K.team_pairs + lg;
        const int nkt = K.team3_kt;                             // a multiple of U (padded with zero columns)
        int pk[U];
        double b[U][TN3];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            pk[u] = __ldg(pr + 4 * u);
#pragma unroll
            for (int t = 0; t < TN3; ++t) b[u][t] = __ldg(t3 + ((size_t)u * 4 * TN3 + t) * 32);
        }
#pragma unroll 1
        for (int k0 = 0; k0 < nkt; k0 += U) {
            const int k1 = (k0 + U < nkt) ? k0 + U : k0;       // the last block re-loads itself
            int pkn[U];
            double bn[U][TN3];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                pkn[u] = __ldg(pr + 4 * (k1 + u));
#pragma unroll
                for (int t = 0; t < TN3; ++t) bn[u][t] = __ldg(t3 + ((size_t)(k1 + u) * 4 * TN3 + t) * 32);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                // a pair entry holds the offsets of its two factors in the exchange buffer, (k >> 2) * 32 + (k & 3), as half-words
                // (computing them here from (k, l) was 40 % of the kernel's instructions: profiles/r02_h)
                const double a = xq[pk[u] & 0xffff] * xq[pk[u] >> 16];
#pragma unroll
                for (int t = 0; t < TN3; ++t) dmma884(a3[u % NACC][t][0], a3[u % NACC][t][1], a, b[u][t]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                pk[u] = pkn[u];
#pragma unroll
                for (int t = 0; t < TN3; ++t) b[u][t] = bn[u][t];
            }
        }
#pragma unroll
        for (int i = 0; i < NRW; ++i) {
            double s_ = a3[0][i / 2][i % 2];
#pragma unroll
            for (int k_ = 1; k_ < NACC; ++k_) s_ += a3[k_][i / 2][i % 2];
            g3[i] = s_;
        }
    }
    double fpart = 0., jd = 0., ka = 0., kah = 0.;
#pragma unroll
    for (int i = 0; i < NRW; ++i) {
        const double y = ax[i / 2][i % 2];
        const double lin_r = lin_t[4 * (NRW * w + i) + lg];
        double g = lin_r + y;
        fpart = fma(lin_r, xo[i], fpart);
        fpart = fma(0.5 * xo[i], y, fpart);
        if (C2) {
            const double t = ax[(NRW + i) / 2][(NRW + i) % 2];
            const double u = a2[i / 2][i % 2];
            g += fma(2. * xo[i], t, u);
            fpart = fma(xo[i] * xo[i], t, fpart);
        }
        if (TS::C3) { g += g3[i]; fpart = fma(xo[i] * (1. / 3.), g3[i], fpart); }     // Euler: sum_j x_j dP3/dx_j = 3 P3
        gn[i] = g;
        jd = fma(g, d_o[i], jd);
        const double pn = fma(dt, g, p[i]), vpn = var[i] * pn;
        ka = fma(pn, vpn, ka);
        kah = fma(vpn, h[i], kah);
    }
    team_sum4(fpart, jd, ka, kah, red, rbuf, bar_id, lane, w);
    double fp = fpart;
    ke = ka;
    if (outside) {
        // PolyModel._fj_bound, poly.py:480-503: J = J0 + outer(sfac, H (x - mu) / beta)
        const double f0 = K.c0 + fpart;
        const double cg = ((f0 - K.f_mu) / K.alpha - jd * rbeta) * rbeta;
#pragma unroll
        for (int i = 0; i < NRW; ++i) gn[i] = fma(cg, h[i], gn[i]);
        fp = (beta * f0 - (beta - K.alpha) * K.f_mu) / K.alpha - K.c0;
        const double cd = dt * cg;
        ke = fma(cd, fma(cd, vh2, 2. * kah), ka);
    }
    lp = K.c0 + fp;
}
