// bfb_post.cu -- the steps on either side of the sampler inside Recipe._sam_step / _pos_step on
// device-resident samples:
//   * importance weights of PostStep: weights = exp(logp - logq), truncated at mean(weights) * n^k_trunc
//     (core/recipe.py:1286-1297);
//   * SystematicResampler.run (utils/misc.py:21-110): the resampled indices are argsort(a)[i_all] -- a full argsort of the
//     previous density values (N up to millions) followed by a gather at the n systematic positions.  The sort is a bitonic
//     network over (key, index) pairs: the strides below 2048 of every stage run inside shared memory, the larger ones as
//     one global pass each; ties are broken by the index, so the result is deterministic.
#include "bfb_common.cuh"
#include <algorithm>

struct PostTmp {
    std::vector<void *> v;
    ~PostTmp() { for (void *p : v) cudaFree(p); }
    int alloc(size_t bytes, void **p) { BFB_CUDA(cudaMalloc(p, bytes ? bytes : 8)); v.push_back(*p); return BFB_OK; }
};

// ---- importance weights ----
static __global__ void __launch_bounds__(256) iw_exp_kernel(const double *__restrict__ logp, const double *__restrict__ logq, int64_t N,
                                                           double *__restrict__ w, double *__restrict__ part)
{
    __shared__ double sm[8];
    double s = 0.;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = exp(logp[i] - logq[i]);
        w[i] = v;
        s += v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.; for (int k = 0; k < 8; ++k) t += sm[k]; part[blockIdx.x] = t; }
}
// one block: total in a fixed order (deterministic), then the cap; a second pass clips
static __global__ void iw_total_kernel(const double *__restrict__ part, int nb, int64_t N, double k_trunc, double *__restrict__ out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.;
        for (int b = 0; b < nb; ++b) t += part[b];
        out[0] = t;                                                    // sum of the weights
        out[1] = (k_trunc < 0.) ? INFINITY : t / (double)N * pow((double)N, k_trunc);     // np.mean(weights) * n_is ** k_trunc
    }
}
static __global__ void __launch_bounds__(256) iw_clip_kernel(const double *__restrict__ w, int64_t N, const double *__restrict__ tot,
                                                            double *__restrict__ wt, double *__restrict__ part)
{
    __shared__ double sm[3][8];
    const double cap = tot[1];
    double s = 0., s2 = 0., mx = 0.;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        double v = w[i];
        v = v < 0. ? 0. : (v > cap ? cap : v);                          // np.clip(weights, 0, cap); NaN stays NaN
        wt[i] = v;
        s += v; s2 = fma(v, v, s2); mx = fmax(mx, v);
    }
    s = warp_sum(s); s2 = warp_sum(s2);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(BFB_FULL, mx, o));
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = s; sm[1][threadIdx.x >> 5] = s2; sm[2][threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0., b = 0., c = 0.;
        for (int k = 0; k < 8; ++k) { a += sm[0][k]; b += sm[1][k]; c = fmax(c, sm[2][k]); }
        part[3 * blockIdx.x] = a; part[3 * blockIdx.x + 1] = b; part[3 * blockIdx.x + 2] = c;
    }
}

extern "C" int bfb_importance_weights(bfb_handle h, const double *logp, const double *logq, int64_t N, double k_trunc,
                                      double *weights, double *weights_trunc, double *stats, int loc)
{
    BFB_REQUIRE(h && logp && logq && N > 0, BFB_ERR_ARG, "bfb_importance_weights: bad arguments");
    BFB_CUDA(cudaSetDevice(h->device));
    PostTmp T;
    int rc;
    const double *dlp = logp, *dlq = logq;
    double *dw = weights, *dwt = weights_trunc;
    if (loc == BFB_HOST) {
        void *a, *b, *c, *d;
        if ((rc = T.alloc(8 * N, &a)) || (rc = T.alloc(8 * N, &b)) || (rc = T.alloc(8 * N, &c)) || (rc = T.alloc(8 * N, &d))) return rc;
        BFB_CUDA(cudaMemcpyAsync(a, logp, 8 * N, cudaMemcpyHostToDevice, h->stream));
        BFB_CUDA(cudaMemcpyAsync(b, logq, 8 * N, cudaMemcpyHostToDevice, h->stream));
        dlp = (double *)a; dlq = (double *)b; dw = (double *)c; dwt = (double *)d;
    } else {
        void *c;
        if (!dw) { if ((rc = T.alloc(8 * N, &c))) return rc; dw = (double *)c; }
        if (!dwt) { if ((rc = T.alloc(8 * N, &c))) return rc; dwt = (double *)c; }
    }
    const int nb = (int)std::min<int64_t>((N + 255) / 256, (int64_t)h->sm_count * 8);
    void *part, *tot;
    if ((rc = T.alloc(8 * 3 * (size_t)nb, &part)) || (rc = T.alloc(16, &tot))) return rc;
    BFB_CUDA(cudaEventRecord(h->ev0, h->stream));
    iw_exp_kernel<<<nb, 256, 0, h->stream>>>(dlp, dlq, N, dw, (double *)part);
    iw_total_kernel<<<1, 32, 0, h->stream>>>((const double *)part, nb, N, k_trunc, (double *)tot);
    iw_clip_kernel<<<nb, 256, 0, h->stream>>>(dw, N, (const double *)tot, dwt, (double *)part);
    h->launches += 3;
    BFB_CUDA(cudaGetLastError());
    BFB_CUDA(cudaEventRecord(h->ev1, h->stream));
    std::vector<double> hp(3 * (size_t)nb);
    double ht[2];
    BFB_CUDA(cudaMemcpyAsync(hp.data(), part, 8 * hp.size(), cudaMemcpyDeviceToHost, h->stream));
    BFB_CUDA(cudaMemcpyAsync(ht, tot, 16, cudaMemcpyDeviceToHost, h->stream));
    if (loc == BFB_HOST) {
        if (weights) BFB_CUDA(cudaMemcpyAsync(weights, dw, 8 * N, cudaMemcpyDeviceToHost, h->stream));
        if (weights_trunc) BFB_CUDA(cudaMemcpyAsync(weights_trunc, dwt, 8 * N, cudaMemcpyDeviceToHost, h->stream));
    }
    BFB_CUDA(cudaStreamSynchronize(h->stream));
    BFB_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    if (stats) {      // sum w | cap | sum w_trunc | sum w_trunc^2 | max w_trunc   (effective sample size = (sum wt)^2 / sum wt^2)
        double a = 0., b = 0., c = 0.;
        for (int k = 0; k < nb; ++k) { a += hp[3 * k]; b += hp[3 * k + 1]; c = std::max(c, hp[3 * k + 2]); }
        stats[0] = ht[0]; stats[1] = ht[1]; stats[2] = a; stats[3] = b; stats[4] = c;
    }
    return BFB_OK;
}

// ---- argsort + gather ----
struct KI { double k; int i; };
__device__ __forceinline__ bool ki_less(double ka, int ia, double kb, int ib) { return ka < kb || (ka == kb && ia < ib); }

static __global__ void sort_init_kernel(const double *__restrict__ a, int64_t N, int64_t Np, double *__restrict__ key, int *__restrict__ idx)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Np; i += (int64_t)gridDim.x * blockDim.x) {
        double v = i < N ? a[i] : INFINITY;
        if (v != v) v = INFINITY;                   // NaN sorts last (np.argsort), before the padding (larger index)
        key[i] = v; idx[i] = (int)i;
    }
}
// one compare-exchange pass with partner distance j of the stage with block size k (global memory)
static __global__ void bitonic_global_kernel(double *__restrict__ key, int *__restrict__ idx, int64_t Np, int64_t j, int64_t k)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < Np / 2; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
        const bool up = (lo & k) == 0;
        const double ka = key[lo], kb = key[hi];
        const int ia = idx[lo], ib = idx[hi];
        if (ki_less(kb, ib, ka, ia) == up) { key[lo] = kb; key[hi] = ka; idx[lo] = ib; idx[hi] = ia; }
    }
}
// all passes with partner distance <= 1024 of the stage with block size k (or, with k_from > 0, all stages k_from..2048 from
// scratch) on one 2048-element tile in shared memory
static __global__ void __launch_bounds__(1024) bitonic_shared_kernel(double *__restrict__ key, int *__restrict__ idx, int64_t k, int full)
{
    __shared__ double sk[2048];
    __shared__ int si[2048];
    const int64_t base = (int64_t)blockIdx.x * 2048;
    for (int e = threadIdx.x; e < 2048; e += 1024) { sk[e] = key[base + e]; si[e] = idx[base + e]; }
    __syncthreads();
    for (int64_t kk = full ? 2 : k; kk <= (full ? 2048 : k); kk <<= 1) {
        for (int j = (int)(kk / 2 > 1024 ? 1024 : kk / 2); j > 0; j >>= 1) {
            const int t = threadIdx.x;
            const int lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
            const bool up = ((base + lo) & kk) == 0;
            const double ka = sk[lo], kb = sk[hi];
            const int ia = si[lo], ib = si[hi];
            if (ki_less(kb, ib, ka, ia) == up) { sk[lo] = kb; sk[hi] = ka; si[lo] = ib; si[hi] = ia; }
            __syncthreads();
        }
    }
    for (int e = threadIdx.x; e < 2048; e += 1024) { key[base + e] = sk[e]; idx[base + e] = si[e]; }
}
static __global__ void gather_kernel(const int *__restrict__ idx, const int64_t *__restrict__ pos, int64_t n, int64_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = idx[pos[i]];
}

// out[i] = argsort(a)[pos[i]], i < n  (pos host or device like a; out like pos).  order_out (optional, N ints, device only when
// loc == BFB_DEVICE): the whole permutation.
extern "C" int bfb_argsort_gather(bfb_handle h, const double *a, int64_t N, const int64_t *pos, int64_t n, int64_t *out, int loc)
{
    BFB_REQUIRE(h && a && N > 0 && N < (1ll << 31) && (n == 0 || (pos && out)), BFB_ERR_ARG, "bfb_argsort_gather: bad arguments");
    BFB_CUDA(cudaSetDevice(h->device));
    PostTmp T;
    int rc;
    int64_t Np = 2048;
    while (Np < N) Np <<= 1;
    const double *da = a;
    const int64_t *dpos = pos;
    int64_t *dout = out;
    void *p;
    if (loc == BFB_HOST) {
        if ((rc = T.alloc(8 * N, &p))) return rc;
        BFB_CUDA(cudaMemcpyAsync(p, a, 8 * N, cudaMemcpyHostToDevice, h->stream));
        da = (const double *)p;
        if (n) {
            for (int64_t i = 0; i < n; ++i) BFB_REQUIRE(pos[i] >= 0 && pos[i] < N, BFB_ERR_ARG, "bfb_argsort_gather: position %lld out of range", (long long)pos[i]);
            if ((rc = T.alloc(8 * n, &p))) return rc;
            BFB_CUDA(cudaMemcpyAsync(p, pos, 8 * n, cudaMemcpyHostToDevice, h->stream));
            dpos = (const int64_t *)p;
            if ((rc = T.alloc(8 * n, &p))) return rc;
            dout = (int64_t *)p;
        }
    }
    void *key, *idx;
    if ((rc = T.alloc(8 * Np, &key)) || (rc = T.alloc(4 * Np, &idx))) return rc;
    const int gb = (int)std::min<int64_t>((Np / 2 + 255) / 256, (int64_t)h->sm_count * 16);
    BFB_CUDA(cudaEventRecord(h->ev0, h->stream));
    sort_init_kernel<<<gb, 256, 0, h->stream>>>(da, N, Np, (double *)key, (int *)idx);
    bitonic_shared_kernel<<<(unsigned)(Np / 2048), 1024, 0, h->stream>>>((double *)key, (int *)idx, 0, 1);
    h->launches += 2;
    for (int64_t k = 4096; k <= Np; k <<= 1) {
        for (int64_t j = k / 2; j > 1024; j >>= 1) {
            bitonic_global_kernel<<<gb, 256, 0, h->stream>>>((double *)key, (int *)idx, Np, j, k);
            h->launches++;
        }
        bitonic_shared_kernel<<<(unsigned)(Np / 2048), 1024, 0, h->stream>>>((double *)key, (int *)idx, k, 0);
        h->launches++;
    }
    if (n) { gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>((const int *)idx, dpos, n, dout); h->launches++; }
    BFB_CUDA(cudaGetLastError());
    BFB_CUDA(cudaEventRecord(h->ev1, h->stream));
    if (loc == BFB_HOST && n) BFB_CUDA(cudaMemcpyAsync(out, dout, 8 * n, cudaMemcpyDeviceToHost, h->stream));
    BFB_CUDA(cudaStreamSynchronize(h->stream));
    BFB_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return BFB_OK;
}
