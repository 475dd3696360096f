// micro-benchmarks: dependent-chain latencies on this GPU (cycles per op, one warp, one lane active or all)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, int n, int active) {
    if ((int)threadIdx.x >= active) return;
    double a = out[0], b = 1.0000001, c = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = a * b + c; a = a * b + c; a = a * b + c; a = a * b + c; }
    long long t1 = clock64();
    out[threadIdx.x + 1] = a; if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / (4.0 * n);
}
__global__ void k_dadd_dmul(double* out, int n) {
    double a = out[0], b = 1.0000001, c = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = __dmul_rn(a, b); a = __dadd_rn(a, c); a = __dmul_rn(a, b); a = __dadd_rn(a, c); }
    long long t1 = clock64();
    out[threadIdx.x + 1] = a; if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / (4.0 * n);
}
__global__ void k_dfma_ilp(double* out, int n) {   // 8 independent chains per thread: throughput per warp
    double a[8]; for (int j = 0; j < 8; ++j) a[j] = out[0] + j;
    double b = 1.0000001, c = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = a[j] * b + c;
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += a[j];
    out[threadIdx.x + 1] = s; if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / (8.0 * n);
}
__global__ void k_lds(double* out, int n) {
    __shared__ unsigned short nxt[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) nxt[i] = (unsigned short)((i * 37 + 11) & 4095);
    __syncthreads();
    unsigned v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { v = nxt[v]; v = nxt[v]; v = nxt[v]; v = nxt[v]; }
    long long t1 = clock64();
    out[threadIdx.x + 1] = v; if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / (4.0 * n);
}
__global__ void k_satom(double* out, int n) {
    __shared__ unsigned ctr[32];
    if (threadIdx.x < 32) ctr[threadIdx.x] = 0;
    __syncthreads();
    unsigned v = 0;
    long long t0 = clock64();
    if (threadIdx.x == 0) for (int i = 0; i < n; ++i) { v = atomicAdd(&ctr[v & 31], 1u); v = atomicCAS(&ctr[v & 31], v, v + 1); }
    long long t1 = clock64();
    out[threadIdx.x + 1] = v; if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / (2.0 * n);
}
__device__ __noinline__ double callee(double a, double b) { return a * b + 1.0; }
__global__ void k_call(double* out, int n) {
    double a = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = callee(a, 1.0000001); a = callee(a, 1.0000002); }
    long long t1 = clock64();
    out[threadIdx.x + 1] = a; if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / (2.0 * n);
}
__global__ void k_lmem(double* out, int n) {
    volatile double buf[32];
    for (int j = 0; j < 32; ++j) buf[j] = out[0] + j;
    int idx = threadIdx.x & 31; double a = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a += buf[idx]; idx = ((int)a + idx + 1) & 31; buf[idx] = a; }
    long long t1 = clock64();
    out[threadIdx.x + 1] = a; if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / (1.0 * n);
}
int main() {
    double* d; cudaMalloc(&d, 8 * 128); cudaMemset(d, 0, 8 * 128); double h;
    auto rd = [&](const char* nm) { cudaDeviceSynchronize(); cudaMemcpy(&h, d + 64, 8, cudaMemcpyDeviceToHost); printf("%-40s %.1f cycles\n", nm, h); };
    for (int rep = 0; rep < 2; ++rep) {
        k_dfma<<<1, 32>>>(d, 20000, 1); rd("DFMA dependent (1 lane)");
        k_dfma<<<1, 32>>>(d, 20000, 32); rd("DFMA dependent (32 lanes)");
        k_dadd_dmul<<<1, 32>>>(d, 20000); rd("DMUL/DADD dependent");
        k_dfma_ilp<<<1, 32>>>(d, 20000); rd("DFMA 8 chains: cycles per DFMA (1 warp)");
        k_dfma_ilp<<<1, 128>>>(d, 20000); rd("DFMA 8 chains: per DFMA per warp (4 warps)");
        k_dfma_ilp<<<1, 512>>>(d, 20000); rd("DFMA 8 chains: per DFMA per warp (16 warps)");
        k_lds<<<1, 32>>>(d, 20000); rd("LDS.U16 dependent chain");
        k_satom<<<1, 32>>>(d, 20000); rd("shared atomicAdd/CAS dependent");
        k_call<<<1, 32>>>(d, 20000); rd("noinline call + DFMA");
        k_lmem<<<1, 32>>>(d, 20000); rd("local mem ld+st dependent iteration");
    }
    return 0;
}
