// Does a second / third / fourth warp on the same SM slow the one-cell-skew recurrence down?  (shared FP64 / shuffle pipes)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o microbench5 microbench5.cu
#include <cstdio>
#include <cuda_runtime.h>
#define NSTEP 8192
// MODE 0: every warp runs the chain.  MODE 1: warp 0 runs the chain, the others poll shared memory in a tight loop.
// MODE 2: warp 0 runs the chain, the others run a long independent FP64 stream (no shuffles).
template <int MODE>
__global__ void k(double* __restrict__ dst, long long* cyc, int slot, int nwarps) {
    __shared__ double sm[4 * 32 * 32];
    __shared__ volatile unsigned flag;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4 * 32 * 32; i += blockDim.x) sm[i] = (i < 1024) ? 1.0 + 1e-6 * i : (i < 3072 ? 0.2 : 0.9);
    if (threadIdx.x == 0) flag = 0;
    __syncthreads();
    if (MODE == 1 && wrp > 0) { unsigned n = 0; while (flag == 0 && n < (1u << 26)) n++; dst[threadIdx.x + 4096] = n; return; }
    if (MODE == 2 && wrp > 0) {
        double x = 1.0 + lane * 1e-3, y = 0.5;
        unsigned n = 0;
        while (flag == 0 && n < (1u << 22)) { for (int i = 0; i < 16; i++) { x = x * 0.999999 + 1e-9; y = y * 1.000001 - 1e-9; } n++; }
        dst[threadIdx.x + 4096] = x + y; return;
    }
    double d = 0.1 + lane * 1e-3, cxp = 0.2, acc = 0.0;
    double* mydst = dst + (size_t)wrp * NSTEP * 32;
    long long t0 = clock64();
    for (int b = 0; b < NSTEP / 8; b++) {
        const int s0 = (b * 8) & 31;
        double a[8], cx[8], cy[8], pr[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int o = ((s0 + u) & 31) * 32 + lane;
            a[u] = sm[o]; cx[u] = sm[1024 + o]; cy[u] = sm[2048 + o]; pr[u] = sm[3072 + o];
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            double up = __shfl_up_sync(0xffffffffu, d, 1);
            if (lane == 0) up = 0.25;
            double v = a[u] - cxp * d;
            v = v - cy[u] * up;
            d = v * pr[u];
            cxp = cx[u];
            mydst[(size_t)(b * 8 + u) * 32 + lane] = d;
            acc += d * a[u];
        }
    }
    long long t1 = clock64();
    if (wrp == 0) flag = 1;
    mydst[lane] = d + acc;
    if (threadIdx.x == 0) cyc[slot] = t1 - t0;
}
int main() {
    double* dst; long long* cyc;
    cudaMalloc(&dst, (size_t)4 * NSTEP * 32 * 8 + 65536); cudaMalloc(&cyc, 256); cudaMemset(cyc, 0, 256);
    int s = 0;
    for (int rep = 0; rep < 2; rep++) {
        s = 0;
        for (int nw = 1; nw <= 4; nw++) k<0><<<1, 32 * nw>>>(dst, cyc, s++, nw);
        for (int nw = 2; nw <= 4; nw++) k<1><<<1, 32 * nw>>>(dst, cyc, s++, nw);
        for (int nw = 2; nw <= 4; nw++) k<2><<<1, 32 * nw>>>(dst, cyc, s++, nw);
    }
    cudaDeviceSynchronize();
    long long c[32]; cudaMemcpy(c, cyc, 256, cudaMemcpyDeviceToHost);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    const char* n[] = {"1 chain warp", "2 chain warps", "3 chain warps", "4 chain warps", "1 chain + 1 smem-polling warp", "1 chain + 2 polling", "1 chain + 3 polling",
                       "1 chain + 1 FP64-stream warp", "1 chain + 2 FP64-stream", "1 chain + 3 FP64-stream"};
    for (int i = 0; i < s; i++) printf("%-34s %.1f cycles/step\n", n[i], c[i] / (double)NSTEP);
    return 0;
}
