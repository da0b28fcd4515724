// What makes a wavefront step slow?  Incrementally adds the pieces of the real sweep step
// to the bare dependency chain.  1 block x NW warps; reports cycles/step for warp 0.
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 2048
template <bool STG, bool IDX, bool LDSOP, bool STVOL>
__global__ void k_step(double* out, double* gdst, double a, double b, long long* cyc, int slot) {
    __shared__ double sm[4][8 * 32 * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = lane; i < 8 * 32 * 4; i += 32) sm[warp][i] = b + 1e-9 * i;
    __syncwarp();
    double d = a, cur = a * lane;
    double* pd = gdst + (size_t)warp * (N_IT + 8) * 32 + lane;
    long long t0 = clock64();
    for (int blk = 0; blk < N_IT / 8; blk++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            double c1 = b, c2 = b, pr = b, aa = a;
            if (LDSOP) { aa = sm[warp][j * 32 + lane]; c1 = sm[warp][256 + j * 32 + lane]; c2 = sm[warp][512 + j * 32 + lane]; pr = sm[warp][768 + j * 32 + lane]; }
            double dn = __shfl_up_sync(0xffffffffu, d, 1);
            if (IDX) { double bv = __shfl_sync(0xffffffffu, cur, j); if (lane == 0) dn = bv; }
            double v = aa - c1 * d;
            v = v - c2 * dn;
            d = v * pr;
            if (STG) { if (STVOL) *(volatile double*)(pd + j * 32) = d; else pd[j * 32] = d; }
        }
        pd += 8 * 32;
    }
    long long t1 = clock64();
    out[threadIdx.x] = d;
    if (threadIdx.x == 0) cyc[slot] = t1 - t0;
}
int main() {
    double *out, *gdst; long long* cyc;
    cudaMalloc(&out, 4096 * 8); cudaMalloc(&gdst, (size_t)4 * (N_IT + 8) * 32 * 8 * 2); cudaMalloc(&cyc, 256); cudaMemset(cyc, 0, 256);
    for (int rep = 0; rep < 2; rep++) {
        k_step<false, false, false, false><<<1, 32>>>(out, gdst, 1.000001, 0.4999, cyc, 0);
        k_step<true, false, false, false><<<1, 32>>>(out, gdst, 1.000001, 0.4999, cyc, 1);
        k_step<false, true, false, false><<<1, 32>>>(out, gdst, 1.000001, 0.4999, cyc, 2);
        k_step<false, false, true, false><<<1, 32>>>(out, gdst, 1.000001, 0.4999, cyc, 3);
        k_step<true, true, true, false><<<1, 32>>>(out, gdst, 1.000001, 0.4999, cyc, 4);
        k_step<true, true, true, false><<<1, 128>>>(out, gdst, 1.000001, 0.4999, cyc, 5);
        k_step<true, false, false, true><<<1, 32>>>(out, gdst, 1.000001, 0.4999, cyc, 6);
        k_step<true, true, true, false><<<8, 128>>>(out, gdst, 1.000001, 0.4999, cyc, 7);
    }
    cudaDeviceSynchronize();
    long long h[32]; cudaMemcpy(h, cyc, 256, cudaMemcpyDeviceToHost);
    const char* names[] = {"bare chain (shfl+DP)", "+STG", "+SHFL.IDX feed", "+LDS operands", "all (1 warp)", "all (4 warps/CTA)", "+STG volatile", "all (8 CTAs x 4 warps, same dst!)"};
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 8; i++) printf("%-36s %.1f cycles/step\n", names[i], h[i] / (double)N_IT);
    return 0;
}
