// Which instruction between the FP64 result and the shuffle costs what?  (B200, 1 warp, dependent chain)
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
template <int MODE>
__global__ void k_step(double* out, double a, double b, double f, long long* cyc, int slot) {
    const int lane = threadIdx.x & 31;
    const int src = (lane - 1) & 31;
    double d = a + lane * 1e-9;
    double feed = f;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N_IT; i++) {
        double dn;
        if (MODE == 0) dn = __shfl_up_sync(0xffffffffu, d, 1);                       // bare
        if (MODE == 1) dn = __shfl_sync(0xffffffffu, d, src);                        // variable-lane rotation
        if (MODE == 2) { double give = (lane == 31) ? feed : d; dn = __shfl_sync(0xffffffffu, give, src); }   // select before
        if (MODE == 3) { dn = __shfl_up_sync(0xffffffffu, d, 1); if (lane == 0) dn = feed; }                  // select after
        if (MODE == 4) {   // integer-pipe free: shuffle hi/lo separately via __shfl on ints built with hiloint2double
            int hi = __double2hiint(d), lo = __double2loint(d);
            hi = __shfl_sync(0xffffffffu, hi, src); lo = __shfl_sync(0xffffffffu, lo, src);
            dn = __hiloint2double(hi, lo);
        }
        double v = a - b * d;
        v = v - b * dn;
        d = v * b;
        feed = feed + 1e-12;
    }
    long long t1 = clock64();
    out[threadIdx.x] = d + feed;
    if (threadIdx.x == 0) cyc[slot] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 256); cudaMemset(cyc, 0, 256);
    for (int rep = 0; rep < 2; rep++) {
        k_step<0><<<1, 32>>>(out, 1.000001, 0.4999, 0.3, cyc, 0);
        k_step<1><<<1, 32>>>(out, 1.000001, 0.4999, 0.3, cyc, 1);
        k_step<2><<<1, 32>>>(out, 1.000001, 0.4999, 0.3, cyc, 2);
        k_step<3><<<1, 32>>>(out, 1.000001, 0.4999, 0.3, cyc, 3);
        k_step<4><<<1, 32>>>(out, 1.000001, 0.4999, 0.3, cyc, 4);
    }
    cudaDeviceSynchronize();
    long long h[32]; cudaMemcpy(h, cyc, 256, cudaMemcpyDeviceToHost);
    const char* names[] = {"shfl_up", "shfl idx (rotation)", "select BEFORE shfl", "select AFTER shfl", "explicit hi/lo"};
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 5; i++) printf("%-24s %.1f cycles/step\n", names[i], h[i] / (double)N_IT);
    return 0;
}
