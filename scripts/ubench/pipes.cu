// Microbenchmark: issue throughput of the instruction mixes considered for the order-independent normal-equation sums.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on one B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
constexpr int ITERS = 4096;
constexpr int NACC = 16;

__global__ void k_ffma(float* out, float a, float b) {
    float acc[NACC];
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fmaf(acc[i], a, b);
    float s = 0; for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imadwide(long long* out, int a, int b) {
    long long acc[NACC];
    int x[NACC];
    for (int i = 0; i < NACC; i++) { acc[i] = threadIdx.x + i; x[i] = a + i + threadIdx.x; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] += (long long)x[i] * (long long)(b + it);
    }
    long long s = 0; for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad32(int* out, int a, int b) {
    int acc[NACC];
    int x[NACC];
    for (int i = 0; i < NACC; i++) { acc[i] = threadIdx.x + i; x[i] = a + i + threadIdx.x; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] += x[i] * (b + it);
    }
    int s = 0; for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_f2i(int* out, float a) {
    int acc[NACC];
    float x[NACC];
    for (int i = 0; i < NACC; i++) { acc[i] = 0; x[i] = a * (threadIdx.x + i); }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) { acc[i] ^= __float2int_rn(x[i]); x[i] = __int_as_float(__float_as_int(x[i]) + acc[i]); }
    }
    int s = 0; for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, double a, double b) {
    double acc[NACC];
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    double s = 0; for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the current scheme: FFMA with magic + 3-input integer add (two products per IADD3)
__global__ void k_cur(unsigned* out, float a, float b) {
    unsigned acc[NACC];
    float x[NACC];
    for (int i = 0; i < NACC; i++) { acc[i] = 0; x[i] = a * (threadIdx.x + i); }
    for (int it = 0; it < ITERS; it++) {
        const float y = b + it;
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] += __float_as_uint(fmaf(x[i], y, 12582912.f)) + __float_as_uint(fmaf(x[i], b, 12582912.f));
    }
    unsigned s = 0; for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: F2I + IMAD.WIDE in the ratio of the proposed scheme (14 conversions per 54 products)
__global__ void k_mix(long long* out, float a, int b) {
    long long acc[27];
    for (int i = 0; i < 27; i++) acc[i] = i;
    float x[7];
    for (int i = 0; i < 7; i++) x[i] = a * (threadIdx.x + i + 1);
    for (int it = 0; it < ITERS / 4; it++) {
        int q[7];
#pragma unroll
        for (int i = 0; i < 7; i++) { q[i] = __float2int_rn(x[i] * (float)(it + b)); }
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = i; j < 7; j++) { acc[t] += (long long)q[i] * (long long)q[j]; t++; }
    }
    long long s = 0; for (int i = 0; i < 27; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// float hi/lo split of each product
__global__ void k_hilo(unsigned* out, float a, float b) {
    unsigned hi[27], lo[27];
    for (int i = 0; i < 27; i++) { hi[i] = 0; lo[i] = 0; }
    float x[7];
    for (int i = 0; i < 7; i++) x[i] = a * (threadIdx.x + i + 1);
    for (int it = 0; it < ITERS / 4; it++) {
        float q[7];
#pragma unroll
        for (int i = 0; i < 7; i++) q[i] = x[i] * (float)(it + b);
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = i; j < 7; j++) {
                const float h = fmaf(q[i], q[j], 12582912.f);
                const float hf = h - 12582912.f;
                const float l = fmaf(q[i], q[j], -hf);
                const float lq = fmaf(l, 1048576.f, 12582912.f);
                hi[t] += __float_as_uint(h); lo[t] += __float_as_uint(lq); t++;
            }
    }
    unsigned s = 0; for (int i = 0; i < 27; i++) s += hi[i] ^ lo[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    void* buf; cudaMalloc(&buf, 148 * 8 * 1024 * 8);
    const int grid = 148 * 4, block = 512;
    const double nthr = (double)grid * block;
    float ms;
    ms = timeit([&] { k_ffma<<<grid, block>>>((float*)buf, 1.0001f, 0.5f); });
    printf("FFMA            %8.3f ms  %7.1f Gop/s\n", ms, nthr * ITERS * NACC / ms * 1e-6);
    ms = timeit([&] { k_imad32<<<grid, block>>>((int*)buf, 3, 5); });
    printf("IMAD32          %8.3f ms  %7.1f Gop/s\n", ms, nthr * ITERS * NACC / ms * 1e-6);
    ms = timeit([&] { k_imadwide<<<grid, block>>>((long long*)buf, 3, 5); });
    printf("IMAD.WIDE acc64 %8.3f ms  %7.1f Gop/s\n", ms, nthr * ITERS * NACC / ms * 1e-6);
    ms = timeit([&] { k_f2i<<<grid, block>>>((int*)buf, 1.5f); });
    printf("F2I (+2 alu)    %8.3f ms  %7.1f Gop/s\n", ms, nthr * ITERS * NACC / ms * 1e-6);
    ms = timeit([&] { k_dfma<<<grid, block>>>((double*)buf, 1.0001, 0.5); });
    printf("DFMA            %8.3f ms  %7.1f Gop/s\n", ms, nthr * ITERS * NACC / ms * 1e-6);
    ms = timeit([&] { k_cur<<<grid, block>>>((unsigned*)buf, 1.5f, 0.5f); });
    printf("cur 2FFMA+IADD3 %8.3f ms  %7.1f Gprod/s\n", ms, nthr * ITERS * NACC * 2 / ms * 1e-6);
    ms = timeit([&] { k_mix<<<grid, block>>>((long long*)buf, 1.5f, 5); });
    printf("mix F2I+IMADW   %8.3f ms  %7.1f Gprod/s\n", ms, nthr * (ITERS / 4) * 27 / ms * 1e-6);
    ms = timeit([&] { k_hilo<<<grid, block>>>((unsigned*)buf, 1.5f, 0.5f); });
    printf("hi/lo float     %8.3f ms  %7.1f Gprod/s\n", ms, nthr * (ITERS / 4) * 27 / ms * 1e-6);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
