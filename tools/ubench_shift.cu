// Microbenchmark: where can funnel shifts run?  SHF (ALU pipe) vs IMAD.HI+IMAD (FMA pipe).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_shift ubench_shift.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t f_shf(uint32_t lo, uint32_t hi, int s) { return __funnelshift_r(lo, hi, s); }
// (lo >> s) | (hi << (32-s)) for s in 1..31, with M = 1u << (32-s): high half of lo*M, low half of hi*M
__device__ __forceinline__ uint32_t f_mad(uint32_t lo, uint32_t hi, uint32_t M) { return __umulhi(lo, M) + hi * M; }

template <int MODE>
__global__ void bench(uint32_t* out, int iters, int s, unsigned long long* cyc) {
    uint32_t a[8], b[8], q[8], acc[4] = {0, 0, 0, 0};
    for (int k = 0; k < 8; k++) { a[k] = threadIdx.x * 2654435761u + k; b[k] = a[k] * 40503u + 7; q[k] = b[k] ^ 0x5bd1e995u; }
    const uint32_t M = 1u << (32 - s);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint32_t x;
                if (MODE == 0) x = f_shf(a[k], b[k], s);
                else if (MODE == 1) x = f_mad(a[k], b[k], M);
                else x = (k & 1) ? f_mad(a[k], b[k], M) : f_shf(a[k], b[k], s);
                acc[r] |= x ^ q[k];
            }
            a[r] += acc[r];      // keep everything live and changing
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] ^ acc[1] ^ acc[2] ^ acc[3];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (unsigned long long)(t1 - t0);
}

int main() {
    uint32_t* out; unsigned long long* cyc;
    cudaMalloc(&out, 148 * 4 * 512 * 4); cudaMallocManaged(&cyc, 8);
    const int iters = 20000;
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int mode = 0; mode < 3; mode++) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) bench<0><<<148, warps * 32>>>(out, iters, 13, cyc);
                if (mode == 1) bench<1><<<148, warps * 32>>>(out, iters, 13, cyc);
                if (mode == 2) bench<2><<<148, warps * 32>>>(out, iters, 13, cyc);
                cudaDeviceSynchronize();
            }
            // per iteration: 32 funnel + 32 xor/or (+4 add)
            printf("warps/SM %2d mode %d (%s): %.1f clk per iteration (32 shifts + 32 LOP3), %.2f clk per warp-iteration per SMSP\n",
                   warps, mode, mode == 0 ? "SHF" : mode == 1 ? "IMAD.HI+IMAD" : "half/half", (double)*cyc / iters,
                   (double)*cyc / iters / (warps / 4.0));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
