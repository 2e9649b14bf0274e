// bfb_micro.cu -- FP64 peak microbenchmarks (roofline denominator; MEASURED_PEAKS.json has no FP64 entry).
#include "bfb_common.cuh"

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int KIND>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a0, double b0)
{
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (double)(threadIdx.x + i);
    double a = a0 + 1e-9 * threadIdx.x, b = b0;
    for (int it = 0; it < iters; ++it) {
        if (KIND == 0 || KIND == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
        }
        if (KIND == 1 || KIND == 2) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) dmma884(acc[i], acc[i + 1], a, b);
        }
    }
    double s = 0.;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

// single-warp-per-scheduler DMMA issue test: NACC independent accumulators, operands from registers (SRC = 0) or one
// 8-byte shared-memory load per DMMA (SRC = 1); reports cycles per DMMA seen by one warp
template <int NACC, int SRC>
__global__ void __launch_bounds__(256) dmma_issue_kernel(double *out, long long *cyc, int iters, double a0, double b0)
{
    __shared__ double tab[32 * 128];
    for (int i = threadIdx.x; i < 32 * 128; i += blockDim.x) tab[i] = b0 + 1e-9 * i;
    __syncthreads();
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = (double)(threadIdx.x + i);
    double a = a0 + 1e-9 * threadIdx.x;
    const double *bp = tab + (threadIdx.x & 31);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        asm volatile("" ::: "memory");
#pragma unroll
        for (int k = 0; k < 7; ++k) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                const double b = SRC ? bp[((k * NACC + i) & 127) * 32] : b0;
                dmma884(acc[i][0], acc[i][1], a, b);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0.;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

extern "C" int bfb_dmma_issue_test(bfb_handle h, int nacc, int src, int warps_per_sm, double *cycles_per_dmma)
{
    BFB_REQUIRE(h && cycles_per_dmma, BFB_ERR_ARG, "bfb_dmma_issue_test: bad arguments");
    BFB_CUDA(cudaSetDevice(h->device));
    double *d; long long *c;
    BFB_CUDA(cudaMalloc(&d, 8));
    BFB_CUDA(cudaMalloc(&c, 8));
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
        if (nacc == 15 && src == 0) dmma_issue_kernel<15, 0><<<h->sm_count, 32 * warps_per_sm, 0, h->stream>>>(d, c, iters, 0.999999, 1e-3);
        else if (nacc == 15) dmma_issue_kernel<15, 1><<<h->sm_count, 32 * warps_per_sm, 0, h->stream>>>(d, c, iters, 0.999999, 1e-3);
        else if (nacc == 8 && src == 0) dmma_issue_kernel<8, 0><<<h->sm_count, 32 * warps_per_sm, 0, h->stream>>>(d, c, iters, 0.999999, 1e-3);
        else if (nacc == 8) dmma_issue_kernel<8, 1><<<h->sm_count, 32 * warps_per_sm, 0, h->stream>>>(d, c, iters, 0.999999, 1e-3);
        else if (nacc == 4 && src == 0) dmma_issue_kernel<4, 0><<<h->sm_count, 32 * warps_per_sm, 0, h->stream>>>(d, c, iters, 0.999999, 1e-3);
        else if (nacc == 2 && src == 0) dmma_issue_kernel<2, 0><<<h->sm_count, 32 * warps_per_sm, 0, h->stream>>>(d, c, iters, 0.999999, 1e-3);
        else dmma_issue_kernel<1, 0><<<h->sm_count, 32 * warps_per_sm, 0, h->stream>>>(d, c, iters, 0.999999, 1e-3);
        h->launches++;
        BFB_CUDA(cudaGetLastError());
        BFB_CUDA(cudaStreamSynchronize(h->stream));
    }
    long long cy;
    BFB_CUDA(cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(c);
    const int na = (nacc == 15 || nacc == 8 || nacc == 4 || nacc == 2) ? nacc : 1;
    *cycles_per_dmma = (double)cy / ((double)iters * 7 * na);
    return BFB_OK;
}

extern "C" int bfb_fp64_peak(bfb_handle h, int kind, double *tflops)
{
    BFB_REQUIRE(h && tflops && kind >= 0 && kind <= 2, BFB_ERR_ARG, "bfb_fp64_peak: bad arguments");
    BFB_CUDA(cudaSetDevice(h->device));
    double *d;
    BFB_CUDA(cudaMalloc(&d, 8));
    const int iters = 4096, blocks = h->sm_count * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        BFB_CUDA(cudaEventRecord(h->ev0, h->stream));
        if (kind == 0) fp64_peak_kernel<0><<<blocks, threads, 0, h->stream>>>(d, iters, 0.999999, 1e-3);
        else if (kind == 1) fp64_peak_kernel<1><<<blocks, threads, 0, h->stream>>>(d, iters, 0.999999, 1e-3);
        else fp64_peak_kernel<2><<<blocks, threads, 0, h->stream>>>(d, iters, 0.999999, 1e-3);
        h->launches++;
        BFB_CUDA(cudaGetLastError());
        BFB_CUDA(cudaEventRecord(h->ev1, h->stream));
        BFB_CUDA(cudaStreamSynchronize(h->stream));
        float ms;
        BFB_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(d);
    // flops per thread-iteration: DFMA 16*2 ; DMMA: 8 mma per warp-iteration * 8*8*4*2 flops / 32 threads = 128
    double per = (kind == 0 ? 32. : kind == 1 ? 128. : 160.);
    *tflops = per * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    return BFB_OK;
}
