// bfb_eval.cuh -- warp-cooperative evaluation of the flattened polynomial and of the surrogate density.
// One warp evaluates one point; lane `lane` owns dimensions j = lane + 32*r, r < NPL.
#pragma once
#include "bfb_common.cuh"

// index of (a<b<c) in the packed cubic-3 order of _poly.pyx:169-177
__device__ __forceinline__ int64_t c3_index(int a, int b, int c, int n)
{
    // triples with first index < a: C(n,3) - C(n-a,3)
    int64_t na = n - a;
    int64_t before_a = ((int64_t)n * (n - 1) * (n - 2) - na * (na - 1) * (na - 2)) / 6;
    // within first index a: pairs (b',c') with a<b'<b : C(n-a-1,2) - C(n-b,2)
    int64_t nb = n - b;
    int64_t before_b = ((na - 1) * (na - 2) - nb * (nb - 1)) / 2;
    return before_a + before_b + (c - b - 1);
}

// the same in 32-bit arithmetic (n <= 128: C(128, 3) = 341376): the 64-bit products and the division by 6 of c3_index cost more
// instructions than the row of coefficients they locate
__device__ __forceinline__ int c3_index32(int a, int b, int c, int n)
{
    const int na = n - a, nb = n - b;
    const int before_a = (n * (n - 1) * (n - 2) - na * (na - 1) * (na - 2)) / 6;
    const int before_b = ((na - 1) * (na - 2) - nb * (nb - 1)) / 2;
    return before_a + before_b + (c - b - 1);
}

// Polynomial value f (all lanes) and Jacobian row J (lane-owned entries) of output `o` at the point whose
// coordinates are staged in shared memory xsm[0..n) (and held per lane in x[]).
// Restates _quadratic_f/_j, _cubic_2_f/_j, _cubic_3_f/_j (modules/_poly.pyx:13-137) and _linear (poly.py:339-352).
template <int NPL>
__device__ __forceinline__ void poly_fg(const DevModel &M, int o, const double *xsm, const double (&x)[NPL],
                                        int lane, double &f, double (&J)[NPL])
{
    const int n = M.n, np = M.np;
    double fpart = 0.;
    const double *lin = M.lin + (size_t)o * np;
#pragma unroll
    for (int r = 0; r < NPL; ++r) {
        J[r] = lin[lane + 32 * r];
        fpart = fma(J[r], x[r], fpart);
    }
    if (M.has_quad) {
        const double *S = M.S + (size_t)o * n * np;
        double y[NPL];
#pragma unroll
        for (int r = 0; r < NPL; ++r) y[r] = 0.;
        for (int k = 0; k < n; ++k) {
            double xk = xsm[k];
#pragma unroll
            for (int r = 0; r < NPL; ++r) y[r] = fma(S[(size_t)k * np + lane + 32 * r], xk, y[r]);
        }
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            J[r] += y[r];
            fpart = fma(0.5 * x[r], y[r], fpart);
        }
    }
    if (M.has_c2) {
        const double *A1T = M.A1T + (size_t)o * n * np;
        const double *A2 = M.A2 + (size_t)o * n * np;
        double t[NPL], u[NPL];
#pragma unroll
        for (int r = 0; r < NPL; ++r) { t[r] = 0.; u[r] = 0.; }
        for (int k = 0; k < n; ++k) {
            double xk = xsm[k];
            double xk2 = xk * xk;
#pragma unroll
            for (int r = 0; r < NPL; ++r) {
                t[r] = fma(A1T[(size_t)k * np + lane + 32 * r], xk, t[r]);
                u[r] = fma(A2[(size_t)k * np + lane + 32 * r], xk2, u[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            J[r] += fma(2. * x[r], t[r], u[r]);
            fpart = fma(x[r] * x[r], t[r], fpart);
        }
    }
    if (M.has_c3) {
        // _cubic_3_f / _cubic_3_j (_poly.pyx:86-137).  The packed coefficients a_jkl (j < k < l, lexicographic) are read in
        // rows (j, k): the lanes own l, so every row is one coalesced load and every coefficient is read twice per
        // evaluation (333 KB at n = 64) instead of once per gradient entry through scattered loads.
        //   pass A (j outer):  dP3/dx_j |_(j smallest) = sum_k x_k sum_l a_jkl x_l   -> one warp reduction per j
        //                      dP3/dx_l |_(l largest)  = sum_(j<k) a_jkl x_j x_k      -> lane-private, no reduction
        //                      P3 = sum_j x_j * (first sum)
        //   pass B (k outer):  dP3/dx_k |_(k middle)   = sum_j x_j sum_l a_jkl x_l   -> one warp reduction per k
        const double *c3 = M.c3 + (size_t)o * M.n_c3;
        // Row (j, k) starts at c3_index(j, k, k + 1): computed arithmetically (a table lookup in front of every row made
        // each row two DEPENDENT L2 round trips), and the rows of a (j, .) / (., k) sweep are loaded RB at a time before any
        // of them is used, so that RB * NPL coalesced loads are in flight per warp instead of one.  The evaluation is
        // bound by the latency of these L2 reads (333 KB of coefficients at n = 64 do not fit next to the tree state).
        constexpr int RB = (NPL <= 2) ? 8 : 4;
        double Gl[NPL], f3 = 0.;
#pragma unroll
        for (int r = 0; r < NPL; ++r) Gl[r] = 0.;
        for (int j = 0; j < n - 2; ++j) {
            const double xj = xsm[j];
            double pj = 0.;
            int off = c3_index32(j, j + 1, j + 2, n);              // row (j, j + 1); row (j, k + 1) follows after n - k - 1 entries
            for (int k0 = j + 1; k0 < n - 1; k0 += RB) {
                double a[RB][NPL];
                int o2 = off;
#pragma unroll
                for (int b = 0; b < RB; ++b) {
                    const int k = k0 + b;
                    const double *rp = c3 + o2 - (k + 1);            // rp[l] = a_jkl
#pragma unroll
                    for (int r = 0; r < NPL; ++r) {
                        const int l = lane + 32 * r;
                        a[b][r] = (k < n - 1 && l > k && l < n) ? __ldg(rp + l) : 0.;
                    }
                    o2 += n - k - 1;
                }
#pragma unroll
                for (int b = 0; b < RB; ++b) {
                    const int k = k0 + b;
                    if (k < n - 1) {
                        const double xk = xsm[k], xjk = xj * xk;
#pragma unroll
                        for (int r = 0; r < NPL; ++r) {
                            pj = fma(xk, a[b][r] * x[r], pj);
                            Gl[r] = fma(a[b][r], xjk, Gl[r]);
                        }
                    }
                }
                off = o2;
            }
            const double tot = warp_sum(pj);
            f3 = fma(xj, tot, f3);
#pragma unroll
            for (int r = 0; r < NPL; ++r) if (j == lane + 32 * r) J[r] += tot;
        }
        for (int k = 1; k < n - 1; ++k) {
            double qk = 0.;
            for (int j0 = 0; j0 < k; j0 += RB) {
                double a[RB][NPL];
#pragma unroll
                for (int b = 0; b < RB; ++b) {
                    const int j = j0 + b;
                    const double *rp = c3 + (j < k ? c3_index32(j, k, k + 1, n) : 0) - (k + 1);
#pragma unroll
                    for (int r = 0; r < NPL; ++r) {
                        const int l = lane + 32 * r;
                        a[b][r] = (j < k && l > k && l < n) ? __ldg(rp + l) : 0.;
                    }
                }
#pragma unroll
                for (int b = 0; b < RB; ++b) {
                    const int j = j0 + b;
                    if (j < k) {
                        const double xj = xsm[j];
#pragma unroll
                        for (int r = 0; r < NPL; ++r) qk = fma(xj, a[b][r] * x[r], qk);
                    }
                }
            }
            const double tot = warp_sum(qk);
#pragma unroll
            for (int r = 0; r < NPL; ++r) if (k == lane + 32 * r) J[r] += tot;
        }
#pragma unroll
        for (int r = 0; r < NPL; ++r) J[r] += Gl[r];
        if (lane == 0) fpart += f3;
    }
    f = M.c0[o] + warp_sum(fpart);
}

// sm[k-major][lane] matrix-vector product y_j = sum_k T[k][j] * v[k], v staged in shared memory
template <int NPL>
__device__ __forceinline__ void matvec(const double *T, int n, int np, const double *vsm, int lane, double (&y)[NPL])
{
#pragma unroll
    for (int r = 0; r < NPL; ++r) y[r] = 0.;
    for (int k = 0; k < n; ++k) {
        double vk = vsm[k];
#pragma unroll
        for (int r = 0; r < NPL; ++r) y[r] = fma(T[(size_t)k * np + lane + 32 * r], vk, y[r]);
    }
}

// PolyModel._fun_and_jac with the radial bound (poly.py:466-503) preceded by the module-level rescale
// and followed by jac / scales_diff (core/module.py:80-85, 221-227), in two steps so that a multi-output model pays for the
// rescale, the Mahalanobis radius and the projection once: module_prepare stages the evaluation point (x inside the
// ellipsoid, its projection x_0 outside) in xsm, module_output evaluates output o there and applies _fj_bound.
// xo: point in the surrogate's (un-rescaled) input space.  xsm, dsm: two per-warp shared buffers of np doubles.
template <int NPL>
struct ModuleCtx {
    bool outside;
    double beta;
    double xe[NPL], d[NPL], Hd[NPL];
};

template <int NPL>
__device__ __forceinline__ void module_prepare(const DevModel &M, const double (&xo)[NPL], int lane,
                                               double *xsm, double *dsm, ModuleCtx<NPL> &K)
{
    const int n = M.n, np = M.np;
    double xs[NPL];
#pragma unroll
    for (int r = 0; r < NPL; ++r) {
        int j = lane + 32 * r;
        xs[r] = M.use_scales ? (xo[r] - M.s0[j]) / M.sdiff[j] : xo[r];
        if (j >= n) xs[r] = 0.;
        K.d[r] = 0.; K.Hd[r] = 0.;
    }
    K.outside = false;
    K.beta = 0.;
    if (M.use_bound) {
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            int j = lane + 32 * r;
            K.d[r] = (j < n) ? xs[r] - M.mu[j] : 0.;
            dsm[j] = K.d[r];
        }
        __syncwarp();
        matvec<NPL>(M.HT, n, np, dsm, lane, K.Hd);
        double part = 0.;
#pragma unroll
        for (int r = 0; r < NPL; ++r) part = fma(K.d[r], K.Hd[r], part);
        K.beta = sqrt(warp_sum(part));
        K.outside = K.beta > M.alpha;
        __syncwarp();
    }
#pragma unroll
    for (int r = 0; r < NPL; ++r) {
        int j = lane + 32 * r;
        // _fj_bound, poly.py:480-503: outside points are evaluated at their projection onto the ellipsoid
        K.xe[r] = !K.outside ? xs[r] : (j < n) ? (M.alpha * xs[r] + (K.beta - M.alpha) * M.mu[j]) / K.beta : 0.;
        xsm[j] = K.xe[r];
    }
    __syncwarp();
}

template <int NPL>
__device__ __forceinline__ void module_output(const DevModel &M, int o, const ModuleCtx<NPL> &K, int lane,
                                              const double *xsm, double &f, double (&J)[NPL])
{
    if (!K.outside) {
        poly_fg<NPL>(M, o, xsm, K.xe, lane, f, J);
    } else {
        const double alpha = M.alpha, beta = K.beta;
        double jj0[NPL], ff0;
        poly_fg<NPL>(M, o, xsm, K.xe, lane, ff0, jj0);
        const double fmu = M.f_mu[o];
        f = (beta * ff0 - (beta - alpha) * fmu) / alpha;
        double part = 0.;
#pragma unroll
        for (int r = 0; r < NPL; ++r) part = fma(jj0[r], K.d[r], part);
        double jd = warp_sum(part);
        double s = (ff0 - fmu) / alpha - jd / beta;
#pragma unroll
        for (int r = 0; r < NPL; ++r) J[r] = jj0[r] + s * (K.Hd[r] / beta);
    }
    if (M.use_scales) {
#pragma unroll
        for (int r = 0; r < NPL; ++r) J[r] = J[r] / M.sdiff[lane + 32 * r];
    }
}

template <int NPL>
__device__ __forceinline__ void module_fg(const DevModel &M, int o, const double (&xo)[NPL], int lane,
                                          double *xsm, double *dsm, double &f, double (&J)[NPL])
{
    ModuleCtx<NPL> K;
    module_prepare<NPL>(M, xo, lane, xsm, dsm, K);
    module_output<NPL>(M, o, K, lane, xsm, f, J);
    __syncwarp();
}

// transforms/_constraint.pyx:133-221 for one coordinate: value, first and second derivative of to_original
__device__ __forceinline__ void to_original_1(double t, double lo, double w, int hb, double &f, double &j, double &jj)
{
    double tf, tj, tjj;
    if (hb == 3) {
        tf = 1. / (1. + exp(-t));
        tj = tf * (1. - tf);
        double e = exp(t);
        tjj = -e * (e - 1.) / (e + 1.) / (e + 1.) / (e + 1.);
    } else if (hb == 1) {
        tf = exp(t); tj = tf; tjj = tf;
    } else if (hb == 2) {
        double e = exp(t);
        tf = 1. - e; tj = -e; tjj = -e;
    } else {
        tf = t; tj = 1.; tjj = 0.;
    }
    f = lo + tf * w; j = tj * w; jj = tjj * w;
}

// Density.logp_and_grad(x, original_space=False, use_surrogate=True) for a surrogate-only density
// (core/density.py:487-566, 724-754).  xt: point in the transformed space. logp on all lanes.
template <int NPL>
__device__ __forceinline__ void density_eval(const DevModel &M, const double (&xt)[NPL], int lane,
                                             double *xsm, double *dsm, double &logp, double (&grad)[NPL])
{
    const int n = M.n, np = M.np;
    double xo[NPL], tj[NPL], tjj[NPL];
#pragma unroll
    for (int r = 0; r < NPL; ++r) {
        int j = lane + 32 * r;
        if (M.use_transform && j < n) to_original_1(xt[r], M.r_lo[j], M.r_w[j], M.hb[j], xo[r], tj[r], tjj[r]);
        else { xo[r] = xt[r]; tj[r] = 1.; tjj[r] = 0.; }
    }
    double f, J[NPL];
    if (M.epilogue) {
        // two-module pipeline (density.py:487-566): Gaussian likelihood of the m pre-whitened surrogate outputs,
        // logp = c0 - 1/2 sum_o f_o^2, Jacobians chained: grad = -sum_o f_o J_o
        double acc = 0., gs[NPL];
#pragma unroll
        for (int r = 0; r < NPL; ++r) gs[r] = 0.;
        ModuleCtx<NPL> K;
        module_prepare<NPL>(M, xo, lane, xsm, dsm, K);
        for (int o = 0; o < M.m; ++o) {
            module_output<NPL>(M, o, K, lane, xsm, f, J);
            acc = fma(f, f, acc);
#pragma unroll
            for (int r = 0; r < NPL; ++r) gs[r] = fma(-f, J[r], gs[r]);
        }
        __syncwarp();
        f = M.e_c0 - 0.5 * acc;
#pragma unroll
        for (int r = 0; r < NPL; ++r) grad[r] = gs[r] * tj[r];
    } else {
        module_fg<NPL>(M, 0, xo, lane, xsm, dsm, f, J);
#pragma unroll
        for (int r = 0; r < NPL; ++r) grad[r] = J[r] * tj[r];
    }
    if (M.use_prior) {
        // third module (inputs ['like', 'x'], examples/des-y1-w-cosmosis.ipynb cells 12, 14): Gaussian prior on the original-space
        // inputs, its Jacobian chained with the variable transform like the surrogate's (density.py:533-565)
        double part = 0.;
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            int j = lane + 32 * r;
            if (j < n) {
                const double dk = xo[r] - M.p_mu[j], wk = M.p_w[j];
                part = fma(wk * dk, dk, part);
                grad[r] -= wk * dk * tj[r];
            }
        }
        f += M.p_c0 - 0.5 * warp_sum(part);
    }
    if (M.use_decay) {
        double d[NPL], Hd[NPL];
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            int j = lane + 32 * r;
            d[r] = (j < n) ? xo[r] - M.d_mu[j] : 0.;
            dsm[j] = d[r];
        }
        __syncwarp();
        matvec<NPL>(M.d_H, n, np, dsm, lane, Hd);
        double part = 0.;
#pragma unroll
        for (int r = 0; r < NPL; ++r) part = fma(d[r], Hd[r], part);
        double beta2 = warp_sum(part);
        __syncwarp();
        double ex = beta2 - M.d_alpha2;
        f -= M.d_gamma * (ex > 0. ? ex : 0.);
        if (beta2 > M.d_alpha2) {
#pragma unroll
            for (int r = 0; r < NPL; ++r) grad[r] -= 2. * M.d_gamma * Hd[r];
        }
    }
    if (M.use_transform) {
        double part = 0.;
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            int j = lane + 32 * r;
            if (j < n) { part += log(fabs(tj[r])); grad[r] += tjj[r] / tj[r]; }
        }
        f += warp_sum(part);
    }
    logp = f;
}
