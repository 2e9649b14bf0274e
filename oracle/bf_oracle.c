/*
 * bf_oracle.c -- CPU ORACLE for bayesfast_b200.  TEST INFRASTRUCTURE ONLY (see bf_oracle.h).
 *
 * Every function restates, in the reference's own evaluation order, the part of h3jia/bayesfast
 * it cites (paths relative to the reference root).  Build with -ffp-contract=off so that the
 * arithmetic matches the reference's Cython (compiled without FMA contraction).
 * Parity pin: tests/golden/*.npz produced from the real reference (tests/golden/make_golden.py).
 */
#include "bf_oracle.h"
#include "../include/bfb_rng.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * bayesfast/modules/_poly.pyx
 * ---------------------------------------------------------------------------------------- */

/* _poly.pyx:13-28 */
void bfo_quadratic_f(const double *x, const double *a, double *out, int m, int n)
{
    for (int i = 0; i < m; ++i) {
        const double *ai = a + (size_t)i * n * n;
        double acc = 0.;
        for (int j = 0; j < n; ++j) {
            double t = 0.;
            for (int k = j; k < n; ++k) t += ai[j * n + k] * x[k];
            acc += t * x[j];
        }
        out[i] = acc;
    }
}

/* _poly.pyx:34-43 */
void bfo_quadratic_j(const double *x, const double *a, double *out, int m, int n)
{
    for (int i = 0; i < m; ++i) {
        const double *ai = a + (size_t)i * n * n;
        for (int j = 0; j < n; ++j) {
            double o = 2 * ai[j * n + j] * x[j];
            for (int k = 0; k < j; ++k) o += ai[k * n + j] * x[k];
            for (int k = j + 1; k < n; ++k) o += ai[j * n + k] * x[k];
            out[i * n + j] = o;
        }
    }
}

/* _poly.pyx:49-64 */
void bfo_cubic_2_f(const double *x, const double *a, double *out, int m, int n)
{
    for (int i = 0; i < m; ++i) {
        const double *ai = a + (size_t)i * n * n;
        double acc = 0.;
        for (int j = 0; j < n; ++j) {
            double t = 0.;
            for (int k = 0; k < n; ++k) t += ai[j * n + k] * x[k];
            acc += t * x[j] * x[j];
        }
        out[i] = acc;
    }
}

/* _poly.pyx:70-80 */
void bfo_cubic_2_j(const double *x, const double *a, double *out, int m, int n)
{
    for (int i = 0; i < m; ++i) {
        const double *ai = a + (size_t)i * n * n;
        for (int j = 0; j < n; ++j) {
            double o = 0.;
            for (int k = 0; k < n; ++k) o += ai[j * n + k] * x[k];
            o *= 2. * x[j];
            for (int k = 0; k < n; ++k) o += ai[k * n + j] * x[k] * x[k];
            out[i * n + j] = o;
        }
    }
}

/* _poly.pyx:86-106 */
void bfo_cubic_3_f(const double *x, const double *a, double *out, int m, int n)
{
    for (int i = 0; i < m; ++i) {
        const double *ai = a + (size_t)i * n * n * n;
        double acc = 0.;
        for (int j = 0; j + 2 < n; ++j) {
            double s = 0.;
            for (int k = j + 1; k + 1 < n; ++k) {
                double t = 0.;
                for (int l = k + 1; l < n; ++l) t += ai[((size_t)j * n + k) * n + l] * x[l];
                s += t * x[k];
            }
            acc += s * x[j];
        }
        out[i] = acc;
    }
}

/* _poly.pyx:112-137 */
void bfo_cubic_3_j(const double *x, const double *a, double *out, int m, int n)
{
    for (int i = 0; i < m; ++i) {
        const double *ai = a + (size_t)i * n * n * n;
        for (int j = 0; j < n; ++j) {
            double o = 0.;
            for (int k = 0; k < j; ++k) {
                double t = 0.;
                for (int l = k + 1; l < j; ++l) t += ai[((size_t)k * n + l) * n + j] * x[l];
                o += t * x[k];
                t = 0.;
                for (int l = j + 1; l < n; ++l) t += ai[((size_t)k * n + j) * n + l] * x[l];
                o += t * x[k];
            }
            for (int k = j + 1; k < n; ++k) {
                double t = 0.;
                for (int l = k + 1; l < n; ++l) t += ai[((size_t)j * n + k) * n + l] * x[l];
                o += t * x[k];
            }
            out[i * n + j] = o;
        }
    }
}

/* _poly.pyx:143-150 */
void bfo_lsq_quadratic(const double *x, double *out, int64_t rows, int n)
{
    int64_t w = (int64_t)n * (n + 1) / 2;
    for (int64_t i = 0; i < rows; ++i) {
        int64_t j = 0;
        for (int k = 0; k < n; ++k)
            for (int l = k; l < n; ++l) out[i * w + j++] = x[i * n + k] * x[i * n + l];
    }
}

/* _poly.pyx:156-163 */
void bfo_lsq_cubic_2(const double *x, double *out, int64_t rows, int n)
{
    int64_t w = (int64_t)n * n;
    for (int64_t i = 0; i < rows; ++i) {
        int64_t j = 0;
        for (int k = 0; k < n; ++k)
            for (int l = 0; l < n; ++l) out[i * w + j++] = x[i * n + k] * x[i * n + k] * x[i * n + l];
    }
}

/* _poly.pyx:169-177 */
void bfo_lsq_cubic_3(const double *x, double *out, int64_t rows, int n)
{
    int64_t w = (int64_t)n * (n - 1) * (n - 2) / 6;
    for (int64_t i = 0; i < rows; ++i) {
        int64_t j = 0;
        for (int k = 0; k < n; ++k)
            for (int l = k + 1; l < n; ++l)
                for (int p = l + 1; p < n; ++p) out[i * w + j++] = x[i * n + k] * x[i * n + l] * x[i * n + p];
    }
}

/* ------------------------------------------------------------------------------------------
 * bayesfast/modules/poly.py
 * ---------------------------------------------------------------------------------------- */

/* poly.py:339-352, 429-441 (_linear, _eval_one) and the scatter-add of :470-478.
 * ff (m) and jj (m*n) must be zeroed by the caller; either may be NULL. */
static void eval_configs(const bfo_poly_model *mod, const double *x, double *ff, double *jj)
{
    int n = mod->n;
    double *xin = (double *)malloc(sizeof(double) * (size_t)n);
    for (int c = 0; c < mod->n_config; ++c) {
        const bfo_config *cf = &mod->configs[c];
        int ni = cf->n_in, no = cf->n_out;
        for (int k = 0; k < ni; ++k) xin[k] = x[cf->in_mask[k]];
        double *f = ff ? (double *)malloc(sizeof(double) * (size_t)no) : NULL;
        double *j = jj ? (double *)malloc(sizeof(double) * (size_t)no * ni) : NULL;
        switch (cf->order) {
        case BFO_LINEAR:
            for (int i = 0; i < no; ++i) {
                const double *ci = cf->coef + (size_t)i * (ni + 1);
                if (f) {
                    double d = 0.;
                    for (int k = 0; k < ni; ++k) d += ci[1 + k] * xin[k];
                    f[i] = d + ci[0];
                }
                if (j) for (int k = 0; k < ni; ++k) j[i * ni + k] = ci[1 + k];
            }
            break;
        case BFO_QUADRATIC:
            if (f) bfo_quadratic_f(xin, cf->coef, f, no, ni);
            if (j) bfo_quadratic_j(xin, cf->coef, j, no, ni);
            break;
        case BFO_CUBIC_2:
            if (f) bfo_cubic_2_f(xin, cf->coef, f, no, ni);
            if (j) bfo_cubic_2_j(xin, cf->coef, j, no, ni);
            break;
        case BFO_CUBIC_3:
            if (f) bfo_cubic_3_f(xin, cf->coef, f, no, ni);
            if (j) bfo_cubic_3_j(xin, cf->coef, j, no, ni);
            break;
        default: break;
        }
        if (f) { for (int i = 0; i < no; ++i) ff[cf->out_mask[i]] += f[i]; free(f); }
        if (j) {
            for (int i = 0; i < no; ++i)
                for (int k = 0; k < ni; ++k) jj[cf->out_mask[i] * n + cf->in_mask[k]] += j[i * ni + k];
            free(j);
        }
    }
    free(xin);
}

/* beta = sqrt((x-mu) H (x-mu)), np.dot(np.dot(x - mu, hess), x - mu)**0.5  (poly.py:467-469, 481) */
static double mahalanobis(const double *x, const double *mu, const double *hess, int n, double *d_out)
{
    double *d = d_out ? d_out : (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) d[i] = x[i] - mu[i];
    double acc = 0.;
    for (int k = 0; k < n; ++k) {
        double t = 0.;
        for (int i = 0; i < n; ++i) t += d[i] * hess[i * n + k];
        acc += t * d[k];
    }
    if (!d_out) free(d);
    return acc; /* squared */
}

/* poly.py:466-478 with :480-503 (_fj_bound) */
void bfo_poly_fun_and_jac(const bfo_poly_model *mod, const double *x, double *f, double *jac)
{
    int n = mod->n, m = mod->m;
    if (mod->use_bound) {
        double *d = (double *)malloc(sizeof(double) * (size_t)n);
        double beta = sqrt(mahalanobis(x, mod->mu, mod->hess, n, d));
        if (beta > mod->alpha) {
            double alpha = mod->alpha;
            double *x0 = (double *)malloc(sizeof(double) * (size_t)n);
            double *ff0 = (double *)calloc((size_t)m, sizeof(double));
            double *jj0 = (double *)calloc((size_t)m * n, sizeof(double));
            double *gb = (double *)malloc(sizeof(double) * (size_t)n);
            for (int i = 0; i < n; ++i) x0[i] = (alpha * x[i] + (beta - alpha) * mod->mu[i]) / beta;
            eval_configs(mod, x0, ff0, jj0);
            for (int i = 0; i < n; ++i) {
                double t = 0.;
                for (int k = 0; k < n; ++k) t += mod->hess[i * n + k] * d[k];
                gb[i] = t / beta;
            }
            for (int o = 0; o < m; ++o) {
                f[o] = (beta * ff0[o] - (beta - alpha) * mod->f_mu[o]) / alpha;
                double jd = 0.;
                for (int k = 0; k < n; ++k) jd += jj0[o * n + k] * d[k];
                double s = (ff0[o] - mod->f_mu[o]) / alpha - jd / beta;
                for (int k = 0; k < n; ++k) jac[o * n + k] = jj0[o * n + k] + s * gb[k];
            }
            free(x0); free(ff0); free(jj0); free(gb); free(d);
            return;
        }
        free(d);
    }
    memset(f, 0, sizeof(double) * (size_t)m);
    memset(jac, 0, sizeof(double) * (size_t)m * n);
    eval_configs(mod, x, f, jac);
}

/* ModuleBase._fun_and_jac_wrapped (core/module.py:221-227) around PolyModel._fun_and_jac */
static void module_fun_and_jac(const bfo_poly_model *mod, const double *x, double *f, double *jac)
{
    int n = mod->n, m = mod->m;
    if (mod->use_scales) {
        double *xs = (double *)malloc(sizeof(double) * (size_t)n);
        for (int i = 0; i < n; ++i) xs[i] = (x[i] - mod->s0[i]) / mod->sdiff[i];
        bfo_poly_fun_and_jac(mod, xs, f, jac);
        for (int o = 0; o < m; ++o)
            for (int k = 0; k < n; ++k) jac[o * n + k] = jac[o * n + k] / mod->sdiff[k];
        free(xs);
    } else {
        bfo_poly_fun_and_jac(mod, x, f, jac);
    }
}

void bfo_poly_eval_batch(const bfo_poly_model *mod, const double *X, int64_t C, double *F, double *J, int n_threads)
{
    int n = mod->n, m = mod->m;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(n_threads)
#endif
    for (int64_t c = 0; c < C; ++c)
        module_fun_and_jac(mod, X + c * n, F + c * m, J + c * (int64_t)m * n);
}

/* ------------------------------------------------------------------------------------------
 * bayesfast/transforms/_constraint.pyx
 * ---------------------------------------------------------------------------------------- */

/* _to_original_f :133-150, _to_original_j :167-186, _to_original_jj :203-221 */
void bfo_to_original(const double *x, const double *ranges, const uint8_t *hb, int n, double *f, double *j, double *jj)
{
    for (int i = 0; i < n; ++i) {
        double lo = ranges[2 * i], hi = ranges[2 * i + 1], w = hi - lo;
        int b0 = hb[2 * i], b1 = hb[2 * i + 1];
        double t = x[i], tf, tj, tjj;
        if (b0 && b1) {
            tf = 1. / (1. + exp(-t));
            tj = 1. / (1. + exp(-t)); tj = tj * (1. - tj);
            double e = exp(t);
            tjj = -e * (e - 1.) / (e + 1.) / (e + 1.) / (e + 1.);
        } else if (b0 && !b1) {
            tf = exp(t); tj = exp(t); tjj = exp(t);
        } else if (!b0 && b1) {
            tf = 1. - exp(t); tj = -exp(t); tjj = -exp(t);
        } else {
            tf = t; tj = 1.; tjj = 0.;
        }
        if (f) f[i] = lo + tf * w;
        if (j) j[i] = tj * w;
        if (jj) jj[i] = tjj * w;
    }
}

/* _from_original_f :19-38; returns 1 + index of the first out-of-bound variable, 0 if fine */
int bfo_from_original(const double *x, const double *ranges, const uint8_t *hb, int n, double *f)
{
    for (int i = 0; i < n; ++i) {
        double lo = ranges[2 * i], hi = ranges[2 * i + 1];
        int b0 = hb[2 * i], b1 = hb[2 * i + 1];
        double t = (x[i] - lo) / (hi - lo);
        if (b0 && b1) {
            if (t <= 0. || t >= 1.) return i + 1;
            t = log(t / (1. - t));
        } else if (b0 && !b1) {
            if (t <= 0.) return i + 1;
            t = log(t);
        } else if (!b0 && b1) {
            if (t >= 1.) return i + 1;
            t = log(1. - t);
        }
        f[i] = t;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * bayesfast/core/density.py: Pipeline.fun_and_jac :487-566 (surrogate-only pipeline) and
 * Density.logp_and_grad :724-754 with original_space=False, use_surrogate=True
 * ---------------------------------------------------------------------------------------- */
void bfo_logp_and_grad(const bfo_density *den, const double *x, double *logp, double *grad)
{
    const bfo_poly_model *mod = den->model;
    int n = mod->n, m = mod->m;
    double *xo = (double *)malloc(sizeof(double) * (size_t)n);
    double *tj = (double *)malloc(sizeof(double) * (size_t)n);
    double *tjj = (double *)malloc(sizeof(double) * (size_t)n);
    double *f = (double *)malloc(sizeof(double) * (size_t)m);
    double *jac = (double *)malloc(sizeof(double) * (size_t)m * n);
    if (den->use_transform) {
        bfo_to_original(x, den->ranges, den->hard_bounds, n, xo, tj, tjj); /* density.py:503-505 */
    } else {
        for (int i = 0; i < n; ++i) { xo[i] = x[i]; tj[i] = 1.; tjj[i] = 0.; }
    }
    module_fun_and_jac(mod, xo, f, jac);
    double lp = f[0];                                            /* density.py:737 */
    if (den->use_epilogue) {
        /* two-module pipeline (density.py:487-566): the surrogate's m outputs feed a Gaussian-likelihood module
         * logp = c - 1/2 (f - d)^T Cinv (f - d), jac = -(Cinv_sym (f - d))^T; Pipeline chains the Jacobians (J_1 . J_0).
         * m = 1, d = 5, Cinv = 4 is f_1 of examples/2d-donut.ipynb. */
        double *r = (double *)malloc(sizeof(double) * (size_t)m);
        double *w = (double *)malloc(sizeof(double) * (size_t)m);
        for (int o = 0; o < m; ++o) r[o] = f[o] - den->e_d[o];
        double q = 0.;
        for (int o = 0; o < m; ++o) {
            double t = 0., ts = 0.;
            for (int p2 = 0; p2 < m; ++p2) {
                t += den->e_cinv[(size_t)o * m + p2] * r[p2];
                ts += 0.5 * (den->e_cinv[(size_t)o * m + p2] + den->e_cinv[(size_t)p2 * m + o]) * r[p2];
            }
            q += r[o] * t;
            w[o] = -ts;
        }
        lp = den->e_c0 - 0.5 * q;
        for (int k = 0; k < n; ++k) {
            double t = 0.;
            for (int o = 0; o < m; ++o) t += w[o] * jac[(size_t)o * n + k];
            grad[k] = t * tj[k];
        }
        free(r); free(w);
    } else
    for (int k = 0; k < n; ++k) grad[k] = jac[k] * tj[k];         /* density.py:558 np.dot(J, diag) */
    if (den->use_prior) {
        /* third module of the pipeline (inputs ['like', 'x']): logp = like + prior(x), Jacobian [1 | d prior / dx] chained
         * with the variable transform like every other module input (density.py:533-565) */
        double s = 0.;
        for (int k = 0; k < n; ++k) {
            const double dk = xo[k] - den->p_mu[k];
            s += den->p_w[k] * dk * dk;
            grad[k] -= den->p_w[k] * dk * tj[k];
        }
        lp += den->p_c0 - 0.5 * s;
    }
    if (den->use_decay) {                                         /* density.py:740-746 */
        double *d = (double *)malloc(sizeof(double) * (size_t)n);
        double beta2 = mahalanobis(xo, den->d_mu, den->d_hess, n, d);
        double ex = beta2 - den->d_alpha2;
        lp -= den->d_gamma * (ex > 0. ? ex : 0.);
        if (beta2 > den->d_alpha2) {
            for (int k = 0; k < n; ++k) {
                double t = 0.;
                for (int i = 0; i < n; ++i) t += d[i] * den->d_hess[i * n + k];
                grad[k] -= 2 * den->d_gamma * t;
            }
        }
        free(d);
    }
    if (den->use_transform) {                                     /* density.py:747-750 */
        double s = 0.;
        for (int k = 0; k < n; ++k) s += log(fabs(tj[k]));
        lp += s;
        for (int k = 0; k < n; ++k) grad[k] += tjj[k] / tj[k];
    }
    *logp = lp;
    free(xo); free(tj); free(tjj); free(f); free(jac);
}

void bfo_logp_and_grad_batch(const bfo_density *den, const double *X, int64_t C, double *logp, double *grad, int n_threads)
{
    int n = den->model->n;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(n_threads)
#endif
    for (int64_t c = 0; c < C; ++c) bfo_logp_and_grad(den, X + c * n, logp + c, grad + c * n);
}

/* ------------------------------------------------------------------------------------------
 * random stream
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const double *ru, *rz; int64_t n_replay;   /* replay views, or NULL */
    uint64_t seed, chain; int64_t t; int overflow;
} rng_t;

static double rng_uniform(rng_t *r)
{
    int64_t t = r->t++;
    if (r->ru) { if (t >= r->n_replay) { r->overflow = 1; return 0.5; } return r->ru[t]; }
    return bfb_draw_uniform(r->seed, r->chain, (uint64_t)t);
}
static double rng_normal(rng_t *r)
{
    int64_t t = r->t++;
    if (r->rz) { if (t >= r->n_replay) { r->overflow = 1; return 0.; } return r->rz[t]; }
    return bfb_draw_normal(r->seed, r->chain, (uint64_t)t);
}

void bfo_rng_fill(uint64_t seed, uint64_t chain, uint64_t t0, int64_t count, double *u, double *z)
{
    for (int64_t i = 0; i < count; ++i) {
        double uu = bfb_draw_uniform(seed, chain, t0 + (uint64_t)i);
        if (u) u[i] = uu;
        if (z) z[i] = bfb_norminv(uu);
    }
}
void bfo_philox_raw(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out4)
{
    bfb_philox_block b = bfb_philox4x32_10(c0, c1, c2, c3, k0, k1);
    for (int i = 0; i < 4; ++i) out4[i] = b.v[i];
}

/* ------------------------------------------------------------------------------------------
 * bayesfast/samplers/hmc_utils: step_size.py, metrics.py, integration.py
 * ---------------------------------------------------------------------------------------- */

/* step_size.py:10-51 */
typedef struct { double log_step, log_bar, target, hbar, k, t0, mu, gamma; int64_t count; int adapt; } dual_avg;

static void da_init(dual_avg *s, double initial_step, double target, double gamma, double k, double t0, int adapt)
{
    s->log_step = log(initial_step); s->log_bar = s->log_step; s->target = target; s->hbar = 0.;
    s->k = k; s->t0 = t0; s->count = 1; s->mu = log(10. * initial_step); s->gamma = gamma; s->adapt = adapt;
}
static double da_current(const dual_avg *s, int warmup) { return warmup ? exp(s->log_step) : exp(s->log_bar); }
static void da_update(dual_avg *s, double accept_stat, int warmup)
{
    if (!warmup || !s->adapt) return;
    double count = (double)s->count;
    double w = 1. / (count + s->t0);
    s->hbar = ((1. - w) * s->hbar + w * (s->target - accept_stat));
    s->log_step = s->mu - s->hbar * sqrt(count) / s->gamma;
    double mk = pow(count, -s->k);
    s->log_bar = mk * s->log_step + (1. - mk) * s->log_bar;
    s->count += 1;
}

/* metrics.py:333-371 _WeightedVariance */
typedef struct { double n_samples; double *mean, *raw_var; } wvar;
static void wvar_init(wvar *w, int n, const double *mean0, const double *var0, double weight)
{
    w->n_samples = weight;
    w->mean = (double *)calloc((size_t)n, sizeof(double));
    w->raw_var = (double *)calloc((size_t)n, sizeof(double));
    if (mean0) memcpy(w->mean, mean0, sizeof(double) * (size_t)n);
    if (var0) for (int i = 0; i < n; ++i) w->raw_var[i] = var0[i];
    for (int i = 0; i < n; ++i) w->raw_var[i] *= w->n_samples;
}
static void wvar_free(wvar *w) { free(w->mean); free(w->raw_var); }
static void wvar_add(wvar *w, int n, const double *x, double weight)
{
    w->n_samples += 1.;
    for (int i = 0; i < n; ++i) {
        double old_diff = x[i] - w->mean[i];
        w->mean[i] += old_diff / w->n_samples;
        double new_diff = x[i] - w->mean[i];
        w->raw_var[i] += weight * old_diff * new_diff;
    }
}

/* metrics.py:374-417 _WeightedCovariance (dense metric) */
typedef struct { double n_samples; double *mean, *raw_cov; } wcov;
static void wcov_init(wcov *w, int n, const double *mean0, const double *cov0, double weight)
{
    w->n_samples = weight;
    w->mean = (double *)calloc((size_t)n, sizeof(double));
    w->raw_cov = (double *)calloc((size_t)n * n, sizeof(double));
    if (mean0) memcpy(w->mean, mean0, sizeof(double) * (size_t)n);
    if (cov0) memcpy(w->raw_cov, cov0, sizeof(double) * (size_t)n * n);
    else for (int i = 0; i < n; ++i) w->raw_cov[(size_t)i * n + i] = 1.;      /* np.eye(nelem) */
    for (size_t i = 0; i < (size_t)n * n; ++i) w->raw_cov[i] *= w->n_samples;
}
static void wcov_free(wcov *w) { free(w->mean); free(w->raw_cov); }
static void wcov_add(wcov *w, int n, const double *x, double weight, double *old_diff, double *new_diff)
{
    w->n_samples += 1.;
    for (int i = 0; i < n; ++i) {
        old_diff[i] = x[i] - w->mean[i];
        w->mean[i] += old_diff[i] / w->n_samples;
        new_diff[i] = x[i] - w->mean[i];
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) w->raw_cov[(size_t)i * n + j] += weight * new_diff[i] * old_diff[j];
}

/* scipy.linalg.cholesky(cov, lower=True) (LAPACK dpotrf, lower triangle of cov only); returns 0 on a
 * non-positive pivot, in which case `chol` is left untouched (metrics.py:282-287 keeps the old factor) */
static int chol_lower(const double *cov, double *chol, double *work, int n)
{
    for (int j = 0; j < n; ++j) {
        for (int i = j; i < n; ++i) {
            double s = cov[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) s -= work[(size_t)i * n + k] * work[(size_t)j * n + k];
            if (i == j) {
                if (!(s > 0.)) return 0;
                work[(size_t)j * n + j] = sqrt(s);
            } else work[(size_t)i * n + j] = s / work[(size_t)j * n + j];
        }
        for (int i = 0; i < j; ++i) work[(size_t)i * n + j] = 0.;
    }
    memcpy(chol, work, sizeof(double) * (size_t)n * n);
    return 1;
}

/* metrics.py:51-91 QuadMetricDiag, :94-132 QuadMetricFull, :135-211 QuadMetricDiagAdapt, :240-330 QuadMetricFullAdapt */
typedef struct {
    int n, adapt, dense; double *var, *std, *inv_std;
    double *cov, *chol, *work, *d_old, *d_new; int chol_error;
    wvar fg, bg; wcov fgc, bgc; int64_t n_samples, previous_update; int adapt_window, update_window, doubling;
} metric_t;

static void metric_init(metric_t *mt, int n, const double *mean0, const double *var0, const bfo_sampler_cfg *cfg)
{
    mt->n = n; mt->adapt = cfg->adapt_metric; mt->dense = cfg->dense_metric; mt->chol_error = 0;
    mt->var = mt->std = mt->inv_std = mt->cov = mt->chol = mt->work = mt->d_old = mt->d_new = NULL;
    if (mt->dense) {
        size_t nn = (size_t)n * n;
        mt->cov = (double *)malloc(sizeof(double) * nn);
        mt->chol = (double *)calloc(nn, sizeof(double));
        mt->work = (double *)malloc(sizeof(double) * nn);
        mt->d_old = (double *)malloc(sizeof(double) * (size_t)n);
        mt->d_new = (double *)malloc(sizeof(double) * (size_t)n);
        memcpy(mt->cov, var0, sizeof(double) * nn);
        if (!chol_lower(mt->cov, mt->chol, mt->work, n)) mt->chol_error = 1;
        if (mt->adapt) {
            wcov_init(&mt->fgc, n, mean0, var0, cfg->initial_weight);
            wcov_init(&mt->bgc, n, NULL, NULL, 10.);   /* _WeightedCovariance(self._n): weight 10, zero mean, identity */
        }
    } else {
        mt->var = (double *)malloc(sizeof(double) * (size_t)n);
        mt->std = (double *)malloc(sizeof(double) * (size_t)n);
        mt->inv_std = (double *)malloc(sizeof(double) * (size_t)n);
        for (int i = 0; i < n; ++i) { mt->var[i] = var0[i]; mt->std[i] = sqrt(var0[i]); mt->inv_std[i] = 1. / mt->std[i]; }
        if (mt->adapt) {
            wvar_init(&mt->fg, n, mean0, var0, cfg->initial_weight);
            wvar_init(&mt->bg, n, NULL, NULL, 10.);   /* _WeightedVariance(self._n): default weight 10, zero mean/var */
        }
    }
    mt->n_samples = 0; mt->previous_update = 0;
    mt->adapt_window = cfg->adapt_window; mt->update_window = cfg->update_window; mt->doubling = cfg->doubling;
}
static void metric_free(metric_t *mt)
{
    free(mt->var); free(mt->std); free(mt->inv_std);
    free(mt->cov); free(mt->chol); free(mt->work); free(mt->d_old); free(mt->d_new);
    if (mt->adapt && !mt->dense) { wvar_free(&mt->fg); wvar_free(&mt->bg); }
    if (mt->adapt && mt->dense) { wcov_free(&mt->fgc); wcov_free(&mt->bgc); }
}
/* metrics.py:73-75 / :113-115 velocity */
static void metric_velocity(const metric_t *mt, const double *p, double *v)
{
    int n = mt->n;
    if (!mt->dense) { for (int i = 0; i < n; ++i) v[i] = mt->var[i] * p[i]; return; }
    for (int i = 0; i < n; ++i) {
        double s = 0.;
        for (int j = 0; j < n; ++j) s += mt->cov[(size_t)i * n + j] * p[j];
        v[i] = s;
    }
}
/* metrics.py:83-86 / :123-127 random: `z` holds the n normals in draw order */
static void metric_random(const metric_t *mt, const double *z, double *p)
{
    int n = mt->n;
    if (!mt->dense) { for (int i = 0; i < n; ++i) p[i] = mt->inv_std[i] * z[i]; return; }
    /* solve_triangular(chol.T, vals): back substitution with U = L^T */
    for (int i = n - 1; i >= 0; --i) {
        double s = z[i];
        for (int j = i + 1; j < n; ++j) s -= mt->chol[(size_t)j * n + i] * p[j];
        p[i] = s / mt->chol[(size_t)i * n + i];
    }
}
/* metrics.py:186-211 / :289-313 */
static void metric_update(metric_t *mt, const double *sample, int warmup)
{
    if (!mt->adapt || !warmup) return;
    int n = mt->n;
    int64_t delta = mt->n_samples - mt->previous_update;
    if (mt->dense) {
        wcov_add(&mt->fgc, n, sample, 1., mt->d_old, mt->d_new);
        wcov_add(&mt->bgc, n, sample, 1., mt->d_old, mt->d_new);
        if ((delta + 1) % mt->update_window == 0) {
            for (size_t i = 0; i < (size_t)n * n; ++i) mt->cov[i] = mt->fgc.raw_cov[i] / mt->fgc.n_samples;
            if (!chol_lower(mt->cov, mt->chol, mt->work, n)) mt->chol_error = 1;
        }
        if (delta >= mt->adapt_window) {
            wcov_free(&mt->fgc);
            mt->fgc = mt->bgc;
            wcov_init(&mt->bgc, n, NULL, NULL, 10.);
            mt->previous_update = mt->n_samples;
            if (mt->doubling) mt->adapt_window *= 2;
        }
        mt->n_samples += 1;
        return;
    }
    wvar_add(&mt->fg, n, sample, 1.);
    wvar_add(&mt->bg, n, sample, 1.);
    if ((delta + 1) % mt->update_window == 0) {
        for (int i = 0; i < n; ++i) {
            mt->var[i] = mt->fg.raw_var[i] / mt->fg.n_samples;
            mt->std[i] = sqrt(mt->var[i]);
            mt->inv_std[i] = 1. / mt->std[i];
        }
    }
    if (delta >= mt->adapt_window) {
        wvar_free(&mt->fg);
        mt->fg = mt->bg;
        wvar_init(&mt->bg, n, NULL, NULL, 10.);
        mt->previous_update = mt->n_samples;
        if (mt->doubling) mt->adapt_window *= 2;
    }
    mt->n_samples += 1;
}

/* integration.py:10 State */
typedef struct { double *q, *p, *v, *g; double energy, logp; double u, vt, weight; } state_t;   /* u, vt (TState.v), weight: integration.py:13 TState */

typedef struct {
    const bfo_density *den; metric_t *mt; int n;
    const bfo_density *den_base; double logxi;   /* tempered samplers (base_hmc.py:220-231): base density, log xi */
    double *arena; size_t arena_states, arena_used;
    int64_t n_eval;
} integ_t;

static state_t new_state(integ_t *ig)
{
    state_t s;
    if (ig->arena_used >= ig->arena_states) { s.q = s.p = s.v = s.g = NULL; s.energy = s.logp = NAN; return s; }
    double *base = ig->arena + ig->arena_used * 4 * (size_t)ig->n;
    ig->arena_used++;
    s.q = base; s.p = base + ig->n; s.v = base + 2 * ig->n; s.g = base + 3 * ig->n;
    s.energy = s.logp = 0.;
    s.u = s.vt = 0.; s.weight = 1.;
    return s;
}

/* integration.py:104-128: inverse temperature, its derivative, the temperature term of the potential and its derivative */
static double t_beta(double u) { return 1. / (1. + exp(-u)); }
static double t_d_beta(double u) { double e = exp(-u); return e / ((1. + e) * (1. + e)); }
static double t_temp_potential(double u) { return u + 2. * log(1. + exp(-u)); }
static double t_d_temp_potential(double u) { double e = exp(u); return (e - 1.) / (e + 1.); }

/* phi = -logp(q), psi = -(logp_base(q) + logxi) (base_hmc.py:228-231) and their gradients (NULL: values only) */
static void t_potentials(integ_t *ig, const double *q, double *phi, double *dphi, double *psi, double *dpsi, double *tmp)
{
    int n = ig->n;
    double lp;
    bfo_logp_and_grad(ig->den, q, &lp, tmp); ig->n_eval++;
    *phi = -lp;
    if (dphi) for (int i = 0; i < n; ++i) dphi[i] = -tmp[i];
    bfo_logp_and_grad(ig->den_base, q, &lp, tmp);
    *psi = -(lp + ig->logxi);
    if (dpsi) for (int i = 0; i < n; ++i) dpsi[i] = -tmp[i];
}

/* the tail both compute_state (integration.py:139-150) and _step (:205-222) share */
static void t_finish(integ_t *ig, state_t *s, double phi, double psi)
{
    int n = ig->n;
    double kin = 0.;
    for (int i = 0; i < n; ++i) kin += s->p[i] * s->v[i];
    kin = 0.5 * kin + s->vt * s->vt / 2.;
    double beta = t_beta(s->u), U = t_temp_potential(s->u);
    double potential = beta * phi + (1. - beta) * psi + U;
    s->energy = kin + potential;
    s->logp = -phi;
    double delta = phi - psi;
    s->weight = delta == 0. ? 1. : delta / expm1(delta);
}

/* integration.py:130-150 TCpuLeapfrogIntegrator.compute_state, Q = (u, q), P = (v, p) */
static state_t t_compute_state(integ_t *ig, const double *q, double u, const double *p, double vt)
{
    int n = ig->n;
    state_t s = new_state(ig);
    memcpy(s.q, q, sizeof(double) * (size_t)n);
    memcpy(s.p, p, sizeof(double) * (size_t)n);
    s.u = u; s.vt = vt;
    double phi, psi;
    t_potentials(ig, s.q, &phi, NULL, &psi, NULL, s.g);
    metric_velocity(ig->mt, s.p, s.v);
    t_finish(ig, &s, phi, psi);
    return s;
}

/* integration.py:152-222 TCpuLeapfrogIntegrator._step: half drift, full kick at the midpoint, half drift */
static state_t t_leapfrog(integ_t *ig, double epsilon, const state_t *st)
{
    int n = ig->n;
    state_t s = new_state(ig);
    double dt = 0.5 * epsilon;
    double u = st->u, vt = st->vt;
    u += vt * dt;
    for (int i = 0; i < n; ++i) s.q[i] = st->q[i] + dt * st->v[i];         /* axpy */
    double phi, psi;
    double *dphi = s.v, *dpsi = s.g;                                       /* scratch until the velocity is recomputed */
    double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
    t_potentials(ig, s.q, &phi, dphi, &psi, dpsi, tmp);
    double beta = t_beta(u), d_beta = t_d_beta(u), dU = t_d_temp_potential(u);
    double d_pot_du = d_beta * (phi - psi) + dU;
    vt += -d_pot_du * epsilon;
    for (int i = 0; i < n; ++i) {
        double d_pot_dq = beta * dphi[i] + (1. - beta) * dpsi[i];
        s.p[i] = st->p[i] + epsilon * (-d_pot_dq);                         /* axpy */
    }
    u += vt * dt;
    metric_velocity(ig->mt, s.p, s.v);
    for (int i = 0; i < n; ++i) s.q[i] = s.q[i] + dt * s.v[i];             /* axpy */
    s.u = u; s.vt = vt;
    t_potentials(ig, s.q, &phi, NULL, &psi, NULL, tmp);
    free(tmp);
    t_finish(ig, &s, phi, psi);
    return s;
}

/* integration.py:28-34 compute_state; metrics.py:73-81 */
static state_t compute_state(integ_t *ig, const double *q, const double *p)
{
    int n = ig->n;
    state_t s = new_state(ig);
    memcpy(s.q, q, sizeof(double) * (size_t)n);
    memcpy(s.p, p, sizeof(double) * (size_t)n);
    bfo_logp_and_grad(ig->den, s.q, &s.logp, s.g); ig->n_eval++;
    double kin = 0.;
    metric_velocity(ig->mt, s.p, s.v);
    for (int i = 0; i < n; ++i) kin += s.p[i] * s.v[i];
    kin *= 0.5;
    s.energy = kin - s.logp;
    return s;
}

/* integration.py:68-95 _step; metrics.py:88-91 velocity_energy */
static state_t leapfrog(integ_t *ig, double epsilon, const state_t *st)
{
    if (ig->den_base) return t_leapfrog(ig, epsilon, st);
    int n = ig->n;
    state_t s = new_state(ig);
    double dt = 0.5 * epsilon;
    for (int i = 0; i < n; ++i) s.p[i] = st->p[i] + dt * st->g[i];         /* axpy */
    metric_velocity(ig->mt, s.p, s.v);
    for (int i = 0; i < n; ++i) s.q[i] = st->q[i] + epsilon * s.v[i];      /* axpy */
    bfo_logp_and_grad(ig->den, s.q, &s.logp, s.g); ig->n_eval++;
    for (int i = 0; i < n; ++i) s.p[i] = s.p[i] + dt * s.g[i];            /* axpy */
    double kin = 0.;
    metric_velocity(ig->mt, s.p, s.v);
    for (int i = 0; i < n; ++i) kin += s.p[i] * s.v[i];
    kin *= 0.5;
    s.energy = kin - s.logp;
    return s;
}

/* ------------------------------------------------------------------------------------------
 * bayesfast/samplers/nuts.py
 * ---------------------------------------------------------------------------------------- */
typedef struct { const double *q; double energy, logp; double u, weight; } proposal_t;   /* u, weight: tnuts.py:12 TProposal */
typedef struct {
    state_t left, right; double *p_sum; proposal_t proposal; double log_size, accept_sum; int64_t n_proposals;
    int valid;
} subtree_t;

typedef struct {
    integ_t *ig; rng_t *rng; int n;
    state_t start, left, right; proposal_t proposal;
    int depth; double log_size, accept_sum; int64_t n_proposals;
    double *p_sum; double max_energy_change, max_change, start_energy, step_size;
    double *vecs; size_t vec_used, vec_cap;  /* scratch p_sum vectors */
    int nan_flag;
} tree_t;

static double np_logaddexp(double a, double b)
{
    if (a == b) return a + 0.6931471805599453;
    double tmp = a - b;
    if (tmp > 0) return a + log1p(exp(-tmp));
    else if (tmp <= 0) return b + log1p(exp(tmp));
    return tmp;
}

/* nuts.py:200-203 */
static int logbern(tree_t *t, double logp)
{
    if (isnan(logp)) t->nan_flag = 1;
    return log(rng_uniform(t->rng)) < logp;
}

static double *new_vec(tree_t *t)
{
    if (t->vec_used >= t->vec_cap) return NULL;
    return t->vecs + (t->vec_used++) * (size_t)t->n;
}

static double dotn(const double *a, const double *b, int n)
{
    double s = 0.;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* nuts.py:105-132 */
static subtree_t single_step(tree_t *t, const state_t *left, double epsilon, int *diverging)
{
    subtree_t st; memset(&st, 0, sizeof(st));
    state_t right = leapfrog(t->ig, epsilon, left);
    double energy_change = right.energy - t->start_energy;
    if (isnan(energy_change)) energy_change = INFINITY;
    if (fabs(energy_change) > fabs(t->max_energy_change)) t->max_energy_change = energy_change;
    if (fabs(energy_change) < t->max_change) {
        double e = exp(-energy_change);
        st.left = right; st.right = right; st.p_sum = right.p;
        st.proposal.q = right.q; st.proposal.energy = right.energy; st.proposal.logp = right.logp;
        st.proposal.u = right.u; st.proposal.weight = right.weight;
        st.log_size = -energy_change; st.accept_sum = e < 1. ? e : 1.; st.n_proposals = 1; st.valid = 1;
        *diverging = 0;
        return st;
    }
    st.log_size = -INFINITY; st.accept_sum = 0.; st.n_proposals = 1; st.valid = 0;
    *diverging = 1;
    return st;
}

/* nuts.py:134-178 */
static subtree_t build_subtree(tree_t *t, const state_t *left, int depth, double epsilon, int *diverging, int *turning)
{
    int n = t->n;
    if (depth == 0) { *turning = 0; return single_step(t, left, epsilon, diverging); }
    /* scratch discipline: the p_sum of the subtree returned from this call lives at slot `mark`
     * (or inside a State for a leaf); everything above is free again on return. */
    size_t mark = t->vec_used;
    subtree_t tree1 = build_subtree(t, left, depth - 1, epsilon, diverging, turning);
    if (*diverging || *turning) return tree1;
    subtree_t tree2 = build_subtree(t, &tree1.right, depth - 1, epsilon, diverging, turning);
    subtree_t tr; memset(&tr, 0, sizeof(tr));
    tr.left = tree1.left; tr.right = tree2.right;
    if (!(*diverging || *turning)) {
        double *p_sum = new_vec(t);
        for (int i = 0; i < n; ++i) p_sum[i] = tree1.p_sum[i] + tree2.p_sum[i];
        int turn = (dotn(p_sum, tr.left.v, n) <= 0) || (dotn(p_sum, tr.right.v, n) <= 0);
        if (depth > 1) {
            double *ps = new_vec(t);
            for (int i = 0; i < n; ++i) ps[i] = tree1.p_sum[i] + tree2.left.p[i];
            int turn1 = (dotn(ps, tree1.left.v, n) <= 0) || (dotn(ps, tree2.left.v, n) <= 0);
            for (int i = 0; i < n; ++i) ps[i] = tree1.right.p[i] + tree2.p_sum[i];
            int turn2 = (dotn(ps, tree1.right.v, n) <= 0) || (dotn(ps, tree2.right.v, n) <= 0);
            turn = turn | turn1 | turn2;
        }
        *turning = turn;
        tr.log_size = np_logaddexp(tree1.log_size, tree2.log_size);
        if (logbern(t, tree2.log_size - tr.log_size)) tr.proposal = tree2.proposal;
        else tr.proposal = tree1.proposal;
        /* compact: move the merged p_sum down to slot `mark` */
        double *dst = t->vecs + mark * (size_t)n;
        memmove(dst, p_sum, sizeof(double) * (size_t)n);
        t->vec_used = mark + 1;
        tr.p_sum = dst;
    } else {
        tr.p_sum = tree1.p_sum; tr.log_size = tree1.log_size; tr.proposal = tree1.proposal;
        /* tree1.p_sum is either a State's p (leaf) or slot `mark`; keep slot `mark` alive */
        t->vec_used = mark + 1;
    }
    tr.accept_sum = tree1.accept_sum + tree2.accept_sum;
    tr.n_proposals = tree1.n_proposals + tree2.n_proposals;
    tr.valid = 1;
    return tr;
}

/* nuts.py:45-103 */
static void tree_extend(tree_t *t, int direction, int *diverging, int *turning)
{
    int n = t->n;
    subtree_t tree;
    state_t lm_begin, lm_end, rm_begin, rm_end; const double *lm_psum, *rm_psum;
    if (direction > 0) {
        tree = build_subtree(t, &t->right, t->depth, t->step_size, diverging, turning);
        lm_begin = t->left; lm_end = t->right; rm_begin = tree.left; rm_end = tree.right;
        lm_psum = t->p_sum; rm_psum = tree.p_sum;
        t->right = tree.right;
    } else {
        tree = build_subtree(t, &t->left, t->depth, -t->step_size, diverging, turning);
        lm_begin = tree.right; lm_end = tree.left; rm_begin = t->left; rm_end = t->right;
        lm_psum = tree.p_sum; rm_psum = t->p_sum;
        t->left = tree.right;
    }
    t->depth += 1;
    t->accept_sum += tree.accept_sum;
    t->n_proposals += tree.n_proposals;
    if (*diverging || *turning) return;

    double size1 = t->log_size, size2 = tree.log_size;
    if (logbern(t, size2 - size1)) t->proposal = tree.proposal;
    t->log_size = np_logaddexp(t->log_size, tree.log_size);

    /* the reference updates self.p_sum in place (:86) BEFORE forming p_sum1/p_sum2 (:94,:97);
     * when direction > 0 leftmost_p_sum aliases self.p_sum, when direction < 0 rightmost_p_sum does. */
    double *ps1 = new_vec(t), *ps2 = new_vec(t);
    for (int i = 0; i < n; ++i) t->p_sum[i] += tree.p_sum[i];
    {
        int turn = (dotn(t->p_sum, t->left.v, n) <= 0) || (dotn(t->p_sum, t->right.v, n) <= 0);
        for (int i = 0; i < n; ++i) ps1[i] = lm_psum[i] + rm_begin.p[i];
        int turn1 = (dotn(ps1, lm_begin.v, n) <= 0) || (dotn(ps1, rm_begin.v, n) <= 0);
        for (int i = 0; i < n; ++i) ps2[i] = lm_end.p[i] + rm_psum[i];
        int turn2 = (dotn(ps2, lm_end.v, n) <= 0) || (dotn(ps2, rm_end.v, n) <= 0);
        *turning = turn | turn1 | turn2;
    }
    t->vec_used -= 2;
}

/* ------------------------------------------------------------------------------------------
 * chain drivers: base_hmc.py:62-85 (astep), :87-172 (run); nuts.py:205-217; hmc.py:16-49
 * ---------------------------------------------------------------------------------------- */
/* tempered samplers (thmc.py, tnuts.py, base_hmc.py:220-262): base density, log xi, u_0 [C], outputs u / weight [C, n_iter] */
typedef struct { const bfo_density *base; double logxi; const double *u0; double *out_u, *out_w; } temper_t;

static int run_chain(const bfo_density *den, const bfo_sampler_cfg *cfg, int is_nuts, int64_t c, int64_t chain_id,
                     const double *x0, double step0, const double *var0, const double *mean0,
                     const double *ru, const double *rz, int64_t n_replay, bfo_run_out *out, const temper_t *tm)
{
    int n = den->model->n;
    int n_iter = cfg->n_iter;
    int status = 0;
    rng_t rng; rng.ru = ru; rng.rz = rz; rng.n_replay = n_replay; rng.seed = cfg->seed; rng.chain = (uint64_t)chain_id;
    rng.t = 0; rng.overflow = 0;
    dual_avg da; da_init(&da, step0, cfg->target_accept, cfg->gamma, cfg->k, cfg->t0, cfg->adapt_step_size);
    metric_t mt; metric_init(&mt, n, mean0, var0, cfg);
    integ_t ig; ig.den = den; ig.mt = &mt; ig.n = n; ig.n_eval = 0;
    ig.den_base = tm ? tm->base : NULL; ig.logxi = tm ? tm->logxi : 0.;
    double u_cur = tm ? tm->u0[c] : 0.;                                            /* base_hmc.py:236-243 */
    int maxd = is_nuts ? cfg->max_treedepth : 0;
    ig.arena_states = ((size_t)1 << maxd) + (size_t)cfg->n_int_step + 8;
    ig.arena = (double *)malloc(sizeof(double) * ig.arena_states * 4 * (size_t)n);
    ig.arena_used = 0;
    size_t vec_cap = (size_t)(maxd + 4) * 2 + 8;
    double *vecs = (double *)malloc(sizeof(double) * vec_cap * (size_t)n);
    double *q = (double *)malloc(sizeof(double) * (size_t)n);
    double *p0 = (double *)malloc(sizeof(double) * (size_t)n);
    double *psum = (double *)malloc(sizeof(double) * (size_t)n);
    double *zz = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(q, x0, sizeof(double) * (size_t)n);

    /* base_hmc.py:42-46 */
    {
        double lp, *g = (double *)malloc(sizeof(double) * (size_t)n);
        bfo_logp_and_grad(den, q, &lp, g);
        int ok = isfinite(lp);
        for (int i = 0; i < n; ++i) ok = ok && isfinite(g[i]);
        free(g);
        if (!ok) { status = 1; goto done; }
    }

    for (int it = 0; it < n_iter; ++it) {
        int warmup = it < cfg->n_warmup;
        ig.arena_used = 0;
        for (int i = 0; i < n; ++i) zz[i] = rng_normal(&rng);
        metric_random(&mt, zz, p0);                                               /* metrics.py:83-86, 123-127 */
        state_t start;
        if (tm) {
            double v0 = rng_normal(&rng);                                         /* base_hmc.py:244-247 */
            start = t_compute_state(&ig, q, u_cur, p0, v0);
        } else start = compute_state(&ig, q, p0);
        if (!isfinite(start.energy)) { status = 2; break; }
        double st_u = 0., st_w = 1.;
        double step_size = da_current(&da, warmup);
        double accept_stat, st_logp, st_energy, st_echange, st_maxe = 0.;
        int st_depth, st_size, diverging = 0;
        const double *end_q;
        if (is_nuts) {
            tree_t t; memset(&t, 0, sizeof(t));
            t.ig = &ig; t.rng = &rng; t.n = n; t.start = start; t.left = start; t.right = start;
            t.proposal.q = start.q; t.proposal.energy = start.energy; t.proposal.logp = start.logp;
            t.proposal.u = start.u; t.proposal.weight = start.weight;
            t.depth = 0; t.log_size = 0.; t.accept_sum = 0.; t.n_proposals = 0;
            memcpy(psum, start.p, sizeof(double) * (size_t)n); t.p_sum = psum;
            t.max_energy_change = 0.; t.max_change = cfg->max_change; t.start_energy = start.energy; t.step_size = step_size;
            t.vecs = vecs; t.vec_used = 0; t.vec_cap = vec_cap;
            int turning = 0;
            for (int d = 0; d < cfg->max_treedepth; ++d) {
                int direction = logbern(&t, log(0.5)) * 2 - 1;                     /* nuts.py:210 */
                t.vec_used = 0;
                tree_extend(&t, direction, &diverging, &turning);
                if (diverging || turning) break;
            }
            if (t.nan_flag) { status = 3; break; }
            accept_stat = t.accept_sum / (double)t.n_proposals;
            st_logp = t.proposal.logp; st_energy = t.proposal.energy; st_depth = t.depth; st_size = (int)t.n_proposals;
            st_echange = t.proposal.energy - start.energy; st_maxe = t.max_energy_change;
            end_q = t.proposal.q;
            st_u = t.proposal.u; st_w = t.proposal.weight;                         /* tnuts.py:22-33 */
        } else {
            state_t state = start;
            for (int s = 0; s < cfg->n_int_step; ++s) state = leapfrog(&ig, step_size, &state);
            double energy_change;
            if (isfinite(state.energy)) {
                energy_change = start.energy - state.energy;
                diverging = fabs(energy_change) > cfg->max_change;
            } else { energy_change = -INFINITY; diverging = 1; }
            double e = exp(energy_change);
            accept_stat = e < 1. ? e : 1.;
            int accepted;
            if (diverging || rng_uniform(&rng) >= accept_stat) { end_q = start.q; accepted = 0; }
            else { end_q = state.q; accepted = 1; }
            st_logp = state.logp; st_energy = state.energy; st_depth = accepted; st_size = cfg->n_int_step;
            st_echange = energy_change;
            st_u = state.u; st_w = state.weight;                                   /* thmc.py:16-27: of the integrated state, accepted or not */
        }
        if (tm) {
            u_cur = st_u;                                                          /* base_hmc.py:237 u0 = stats._u[-1] */
            if (tm->out_u) tm->out_u[(size_t)c * n_iter + it] = st_u;
            if (tm->out_w) tm->out_w[(size_t)c * n_iter + it] = st_w;
        }
        da_update(&da, accept_stat, warmup);
        metric_update(&mt, end_q, warmup);
        memmove(q, end_q, sizeof(double) * (size_t)n);
        size_t o = (size_t)c * n_iter + it;
        if (out->samples) memcpy(out->samples + o * n, q, sizeof(double) * (size_t)n);
        if (out->logp) out->logp[o] = st_logp;
        if (out->energy) out->energy[o] = st_energy;
        if (out->tree_depth) out->tree_depth[o] = st_depth;
        if (out->tree_size) out->tree_size[o] = st_size;
        if (out->mean_tree_accept) out->mean_tree_accept[o] = accept_stat;
        if (out->step_size) out->step_size[o] = exp(da.log_step);
        if (out->step_size_bar) out->step_size_bar[o] = exp(da.log_bar);
        if (out->energy_change) out->energy_change[o] = st_echange;
        if (out->max_energy_change) out->max_energy_change[o] = st_maxe;
        if (out->diverging) out->diverging[o] = diverging;
    }
done:
    if (rng.overflow && status == 0) status = 4;
    if (out->final_step) {
        out->final_step[c * 4 + 0] = da.log_step; out->final_step[c * 4 + 1] = da.log_bar;
        out->final_step[c * 4 + 2] = da.hbar; out->final_step[c * 4 + 3] = (double)da.count;
    }
    if (out->final_var) {
        if (mt.dense) memcpy(out->final_var + (size_t)c * n * n, mt.cov, sizeof(double) * (size_t)n * n);
        else memcpy(out->final_var + c * n, mt.var, sizeof(double) * (size_t)n);
    }
    if (out->n_draws) out->n_draws[c] = rng.t;
    if (out->status) out->status[c] = status;
    free(ig.arena); free(vecs); free(q); free(p0); free(psum); free(zz);
    metric_free(&mt);
    return status;
}

static int run_all(const bfo_density *den, const bfo_sampler_cfg *cfg, int is_nuts, int64_t C, int64_t chain0,
                   const double *x0, const double *step0, const double *var0, const double *mean0,
                   const double *ru, const double *rz, int64_t n_replay, bfo_run_out *out, const temper_t *tm)
{
    int n = den->model->n;
    int bad = 0;
    int nt = cfg->n_threads;
#ifdef _OPENMP
    if (nt <= 0) nt = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt) reduction(| : bad)
#endif
    for (int64_t c = 0; c < C; ++c) {
        int s = run_chain(den, cfg, is_nuts, c, chain0 + c, x0 + c * n, step0[c], var0 + c * (cfg->dense_metric ? (int64_t)n * n : n), mean0 + c * n,
                          ru ? ru + c * n_replay : NULL, rz ? rz + c * n_replay : NULL, n_replay, out, tm);
        bad |= (s != 0);
    }
    (void)nt;
    return bad;
}

int bfo_nuts_run(const bfo_density *den, const bfo_sampler_cfg *cfg, int64_t C, int64_t chain0,
                 const double *x0, const double *step0, const double *var0, const double *mean0,
                 const double *draws_u, const double *draws_z, int64_t n_replay, bfo_run_out *out)
{
    return run_all(den, cfg, 1, C, chain0, x0, step0, var0, mean0, draws_u, draws_z, n_replay, out, NULL);
}

int bfo_hmc_run(const bfo_density *den, const bfo_sampler_cfg *cfg, int64_t C, int64_t chain0,
                const double *x0, const double *step0, const double *var0, const double *mean0,
                const double *draws_u, const double *draws_z, int64_t n_replay, bfo_run_out *out)
{
    return run_all(den, cfg, 0, C, chain0, x0, step0, var0, mean0, draws_u, draws_z, n_replay, out, NULL);
}

/* TNUTS / THMC (samplers/tnuts.py, thmc.py, hmc_utils/base_hmc.py:220-262, integration.py:98-222); diagonal metric only */
int bfo_tempered_run(const bfo_density *den, const bfo_density *den_base, double logxi, int is_nuts,
                     const bfo_sampler_cfg *cfg, int64_t C, int64_t chain0,
                     const double *x0, const double *u0, const double *step0, const double *var0, const double *mean0,
                     const double *draws_u, const double *draws_z, int64_t n_replay, bfo_run_out *out,
                     double *out_u, double *out_weight)
{
    temper_t tm; tm.base = den_base; tm.logxi = logxi; tm.u0 = u0; tm.out_u = out_u; tm.out_w = out_weight;
    if (cfg->dense_metric) return -1;
    return run_all(den, cfg, is_nuts, C, chain0, x0, step0, var0, mean0, draws_u, draws_z, n_replay, out, &tm);
}
