// Latency microbenchmarks that size the wavefront kernels (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096

__global__ void k_dp_chain(double* out, double a, double b, long long* cyc) {
    double v = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; i++) { v = v * b; v = v - a; }
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dp_sweepstep(double* out, double a, double b, long long* cyc) {
    // the forward-sweep dependency: d = ((A - cx*dprev) - cy*shfl_up(dprev)) * pr
    double d = a;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N_IT; i++) {
        double dn = __shfl_up_sync(0xffffffffu, d, 1);
        double v = a - b * d;
        v = v - b * dn;
        d = v * b;
    }
    long long t1 = clock64();
    out[threadIdx.x] = d;
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
}
__global__ void k_shfl_chain(double* out, double a, long long* cyc) {
    double v = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; i++) v = __shfl_up_sync(0xffffffffu, v, 1);
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
}
__global__ void k_chase(const unsigned* next, unsigned start, int n, unsigned* sink, long long* cyc, int slot) {
    unsigned p = start;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) p = __ldcg(next + p);
    long long t1 = clock64();
    *sink = p;
    cyc[slot] = t1 - t0;
}
__global__ void k_chase_volatile(const unsigned* next, unsigned start, int n, unsigned* sink, long long* cyc, int slot) {
    unsigned p = start;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) p = *(const volatile unsigned*)(next + p);
    long long t1 = clock64();
    *sink = p;
    cyc[slot] = t1 - t0;
}
// ping-pong between two blocks (different SMs) through global memory, flag-in-data
__global__ void k_pingpong(volatile unsigned long long* a, volatile unsigned long long* b, int n, long long* cyc) {
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 1; i <= n; i++) { *a = i; while (*b != (unsigned long long)i) {} }
        long long t1 = clock64();
        cyc[8] = t1 - t0;
    } else {
        for (int i = 1; i <= n; i++) { while (*a != (unsigned long long)i) {} *b = i; }
    }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 128); cudaMemset(cyc, 0, 128);
    k_dp_chain<<<1, 32>>>(out, 1.000001, 0.999999, cyc);
    k_dp_sweepstep<<<1, 32>>>(out, 1.000001, 0.4999, cyc);
    k_shfl_chain<<<1, 32>>>(out, 1.5, cyc);
    // pointer chase: stride 4 KB+ permutation over a buffer
    const int CH = 2000;
    size_t sizes[2] = {32u << 20, 1024u << 20};
    for (int s = 0; s < 2; s++) {
        size_t n = sizes[s] / 4; unsigned* h = (unsigned*)malloc(n * 4);
        size_t stride = 1031 * 64;  // elements; co-prime-ish walk
        for (size_t i = 0; i < n; i++) h[i] = (unsigned)((i + stride) % n);
        unsigned* d; cudaMalloc(&d, n * 4); cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
        unsigned* sink; cudaMalloc(&sink, 4);
        if (s == 0) { // warm L2 by touching the chain twice
            k_chase<<<1, 1>>>(d, 0, CH, sink, cyc, 7); k_chase<<<1, 1>>>(d, 0, CH, sink, cyc, 7);
            k_chase<<<1, 1>>>(d, 0, CH, sink, cyc, 3);
            k_chase_volatile<<<1, 1>>>(d, 0, CH, sink, cyc, 5);
        } else {
            k_chase<<<1, 1>>>(d, 0, CH, sink, cyc, 4);
        }
        cudaDeviceSynchronize(); free(h);
    }
    unsigned long long* flags; cudaMalloc(&flags, 1024); cudaMemset(flags, 0, 1024);
    k_pingpong<<<2, 32>>>(flags, flags + 64, 1000, cyc);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, 128, cudaMemcpyDeviceToHost);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SM clock attr %d kHz; err=%s\n", clk, cudaGetErrorString(cudaGetLastError()));
    printf("DP dependent op (DMUL+DADD pair / 2): %.1f cycles\n", h[0] / (2.0 * N_IT));
    printf("sweep step (shfl + 3 DMUL + 2 DADD, dependent): %.1f cycles\n", h[1] / (double)N_IT);
    printf("64-bit shfl_up chain: %.1f cycles\n", h[2] / (double)N_IT);
    printf("L2-hit load (ld.cg chase, 32 MB): %.1f cycles\n", h[3] / (double)CH);
    printf("L2-hit volatile load: %.1f cycles\n", h[5] / (double)CH);
    printf("DRAM load (chase over 1 GB): %.1f cycles\n", h[4] / (double)CH);
    printf("global ping-pong round trip (2 SMs): %.1f cycles\n", h[8] / 1000.0);
    return 0;
}
