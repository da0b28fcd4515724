// How fast can ONE warp run the forward-sweep recurrence with a one-cell skew (lane l lags lane l-1 by one cell)?
// Models the lean loop of the SKB=1 sweep: operands of U steps preloaded from shared memory, the up-neighbour by
// shuffle, lane-0 edge select, result stored to global memory, edge lane stores to a ring.  B200, 1 warp per block.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o microbench4 microbench4.cu
#include <cstdio>
#include <cuda_runtime.h>
#define NSTEP 8192
template <int U, int MODE>
__global__ void k_lean(double* __restrict__ dst, const double* __restrict__ ops, long long* cyc, int slot) {
    extern __shared__ double sm[];          // [4 arrays][256 steps][32 lanes] ring (re-used cyclically)
    __shared__ double ring[256];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4 * 128 * 32; i += blockDim.x) sm[i] = ops[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) ring[i] = 0.25 + 1e-6 * i;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    double d = 0.1 + lane * 1e-3, cxp = 0.2;
    double acc = 0.0;
    long long t0 = clock64();
    for (int b = 0; b < NSTEP / U; b++) {
        const int s0 = (b * U) & 255;
        double a[U], cx[U], cy[U], pr[U], e[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int o = ((s0 + u) & 127) * 32 + lane;
            a[u] = sm[o]; cx[u] = sm[4096 + o]; cy[u] = sm[8192 + o]; pr[u] = sm[12288 + o];
            e[u] = ring[(s0 + u) & 255];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            double up = __shfl_up_sync(0xffffffffu, d, 1);
            if (lane == 0) up = e[u];
            double v = a[u] - cxp * d;
            v = v - cy[u] * up;
            d = v * pr[u];
            cxp = cx[u];
            if (MODE >= 1) dst[(size_t)(b * U + u) * 32 + lane] = d;
            if (MODE >= 2) { double* p = (lane == 31) ? &ring[(s0 + u + 128) & 255] : &sm[16384 + lane]; *p = d; }
            if (MODE >= 3) acc += d * a[u];
        }
    }
    long long t1 = clock64();
    dst[lane] = d + acc;
    if (lane == 0) cyc[slot] = t1 - t0;
}
int main() {
    double *dst, *ops; long long* cyc;
    cudaMalloc(&dst, (size_t)NSTEP * 32 * 8 + 4096); cudaMalloc(&ops, 4 * 128 * 32 * 8); cudaMalloc(&cyc, 256);
    double* h = (double*)malloc(4 * 128 * 32 * 8);
    for (int i = 0; i < 4 * 128 * 32; i++) h[i] = (i < 4096) ? 1.0 + 1e-6 * i : (i < 12288 ? 0.2 : 0.9);
    cudaMemcpy(ops, h, 4 * 128 * 32 * 8, cudaMemcpyHostToDevice);
    const size_t smem = (4 * 128 * 32 + 64) * 8;
#define RUN(U, M, s) cudaFuncSetAttribute(k_lean<U, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    k_lean<U, M><<<1, 64, smem>>>(dst, ops, cyc, s); k_lean<U, M><<<1, 64, smem>>>(dst, ops, cyc, s);
    RUN(8, 0, 0) RUN(8, 1, 1) RUN(8, 2, 2) RUN(8, 3, 3) RUN(16, 2, 4) RUN(4, 2, 5) RUN(16, 3, 6)
    cudaDeviceSynchronize();
    long long c[32]; cudaMemcpy(c, cyc, 256, cudaMemcpyDeviceToHost);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    const char* n[] = {"U=8 compute only", "U=8 + global store", "U=8 + store + edge ring store", "U=8 + store + ring + dot", "U=16 + store + ring", "U=4 + store + ring", "U=16 + store + ring + dot"};
    for (int i = 0; i < 7; i++) printf("%-34s %.1f cycles/step\n", n[i], c[i] / (double)NSTEP);
    return 0;
}
