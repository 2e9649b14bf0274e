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
// Coefficient operand: host-built table bfrag[kt][tile][lane] (bfb_build_frag, bfb_model.cu), tiles grouped in blocks by
// the A operand they multiply:  [S | A] . x   |   A^T . x^2   |   H . (x - mu).
#pragma once
#include "bfb_common.cuh"

template <int NR, bool C2>
struct DmmaShape {
    static constexpr int TX = C2 ? NR : (NR + 1) / 2;     // tiles of the block multiplying x
    static constexpr int T2 = C2 ? (NR + 1) / 2 : 0;      // tiles multiplying x^2
    static constexpr int TD = (NR + 1) / 2;               // tiles multiplying x - mu
    static constexpr int NT = TX + T2 + TD;
    static constexpr int FRAG_DOUBLES = NR * NT * 32;
};

inline int bfb_frag_tiles(int nr, bool c2) { return (c2 ? nr : (nr + 1) / 2) + (c2 ? (nr + 1) / 2 : 0) + (nr + 1) / 2; }
// instantiated dims-per-lane for input_size n (0: not supported)
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

// lane partials of the polynomial part: gg = lin + S x + 2 x.(A x) + A^T x^2, hh = H (x - mu),
// fpart = sum_own lin x + x.(Sx)/2 + x^2 (A x), bpart = sum_own (x - mu) hh
template <int NR, bool C2>
__device__ __forceinline__ void dmma_core(const double *bsm, int lane, const double (&x)[NR],
                                          const double *mu_t, const double *lin_t, double (&gg)[NR],
                                          double (&hh)[NR], double &fpart, double &bpart)
{
    // mu_t / lin_t: [32] tables indexed by dimension (shared memory); read where used instead of living in registers
    const int lg_ = lane & 3;
    using SH = DmmaShape<NR, C2>;
    double acc[SH::NT][2];
#pragma unroll
    for (int t = 0; t < SH::NT; ++t) acc[t][0] = acc[t][1] = 0.;
    const double *bp = bsm + lane;
    // the table is loop invariant for the callers; without this fence the compiler hoists all B fragments into
    // registers (and spills them) instead of streaming them from shared memory next to the MMAs
    asm volatile("" ::: "memory");
#pragma unroll
    for (int kt = 0; kt < NR; ++kt) {
        const double ax = x[kt], ax2 = x[kt] * x[kt], ad = x[kt] - mu_t[4 * kt + lg_];
#pragma unroll
        for (int t = 0; t < SH::NT; ++t) {
            const double b = bp[(kt * SH::NT + t) * 32];
            const double a = (t < SH::TX) ? ax : (t < SH::TX + SH::T2) ? ax2 : ad;
            dmma884(acc[t][0], acc[t][1], a, b);
        }
    }
    fpart = 0.; bpart = 0.;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const double y = acc[r / 2][r % 2];
        const double h = acc[SH::TX + SH::T2 + r / 2][r % 2];
        const double lin_r = lin_t[4 * r + lg_];
        double g = lin_r + y;
        fpart = fma(lin_r, x[r], fpart);
        fpart = fma(0.5 * x[r], y, fpart);
        if (C2) {
            const double t = acc[(NR + r) / 2][(NR + r) % 2];
            const double u = acc[SH::TX + r / 2][r % 2];
            g += fma(2. * x[r], t, u);
            fpart = fma(x[r] * x[r], t, fpart);
        }
        gg[r] = g; hh[r] = h;
        bpart = fma(x[r] - mu_t[4 * r + lg_], h, bpart);
    }
}

struct DmmaConsts {
    double c0, alpha, alpha2, f_mu;
    int n;
};

// PolyModel._fun_and_jac with the radial bound (poly.py:443-503) for the 8 points of the warp.
// `ke_of` maps the gradient (own dims) to a lane partial that is reduced together with the others (the sampler's
// kinetic energy of the new momentum); `live` masks points whose outside test should not trigger the second pass.
template <int NR, bool C2, class KE>
__device__ __forceinline__ void dmma_logp_grad(const double *bsm, int lane, const DmmaConsts &K,
                                               const double (&x_in)[NR], const double *mu, const double *lin,
                                               bool live, double &lp, double (&gn)[NR], KE &&ke_of, double &ke)
{
    const int lg = lane & 3;
    double x[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) x[r] = x_in[r];
    double gg[NR], hh[NR], fpart, bpart;
    dmma_core<NR, C2>(bsm, lane, x, mu, lin, gg, hh, fpart, bpart);
    double kp = ke_of(gg), zz = 0.;
    qsum4(kp, bpart, fpart, zz, lane);
    double fp = fpart;
    ke = kp;
#pragma unroll
    for (int r = 0; r < NR; ++r) gn[r] = gg[r];
    const bool outside = live && (bpart > K.alpha2);
    if (__any_sync(BFB_FULL, outside)) {
        // PolyModel._fj_bound, poly.py:480-503: project onto the ellipsoid and evaluate there
        // divisions by beta are done as multiplications by 1 / beta: the padded dimensions hold exact zeros and a zero
        // numerator sends the FP64 division to its slow path (measured: one ~100-instruction subroutine call per round)
        const double beta = sqrt(bpart), rbeta = 1. / beta;
        double d0[NR], hd0[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const double mu_r = mu[4 * r + lg];
            d0[r] = x[r] - mu_r; hd0[r] = hh[r];
            if (outside) x[r] = (4 * r + lg < K.n) ? (K.alpha * x[r] + (beta - K.alpha) * mu_r) * rbeta : 0.;
        }
        double f1, b1;
        dmma_core<NR, C2>(bsm, lane, x, mu, lin, gg, hh, f1, b1);
        double jd = 0.;
#pragma unroll
        for (int r = 0; r < NR; ++r) jd = fma(gg[r], d0[r], jd);
        double z1 = 0., z2 = 0.;
        qsum4(f1, jd, z1, z2, lane);
        double g2[NR];
        const double f0 = K.c0 + f1;
        const double sfac = (f0 - K.f_mu) / K.alpha - jd * rbeta;
#pragma unroll
        for (int r = 0; r < NR; ++r) g2[r] = gg[r] + sfac * (hd0[r] * rbeta);
        const double k2 = qsum(ke_of(g2));
        if (outside) {
#pragma unroll
            for (int r = 0; r < NR; ++r) gn[r] = g2[r];
            fp = (beta * f0 - (beta - K.alpha) * K.f_mu) / K.alpha - K.c0;
            ke = k2;
        }
    }
    lp = K.c0 + fp;
}
