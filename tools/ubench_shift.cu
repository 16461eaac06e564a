// Microbenchmark: where can funnel shifts run?  SHF (ALU pipe) vs IMAD.HI+IMAD (FMA pipe).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_shift ubench_shift.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t f_shf(uint32_t lo, uint32_t hi, int s) { return __funnelshift_r(lo, hi, s); }
// (lo >> s) | (hi << (32-s)) for s in 1..31, with M = 1u << (32-s): high half of lo*M, low half of hi*M
__device__ __forceinline__ uint32_t f_mad(uint32_t lo, uint32_t hi, uint32_t M) { return __umulhi(lo, M) + hi * M; }

// MODE 0: 32 SHF + 32 LOP3 ; 1: 32 (IMAD.HI+IMAD) + 32 LOP3 ; 2: 16 SHF + 16 IMAD pairs + 32 LOP3 ; 3: 64 LOP3 only
// 4: 8 of 32 shifts as IMAD pairs
template <int MODE>
__global__ void bench(uint32_t* out, int iters, int s, unsigned long long* cyc) {
    uint32_t a[32], acc[4];
    for (int k = 0; k < 32; k++) { a[k] = threadIdx.x * 2654435761u + k * 40503u; }
    for (int k = 0; k < 4; k++) { acc[k] = k; }
    const uint32_t M = 1u << (32 - s);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 32; k++) {
            uint32_t x;
            const uint32_t lo = a[k], hi = a[(k + 1) & 31];
            bool mad = MODE == 1 || (MODE == 2 && (k & 1)) || (MODE == 4 && (k & 3) == 3);
            if (MODE == 3) x = lo ^ hi;
            else if (mad) x = f_mad(lo, hi, M);
            else x = f_shf(lo, hi, s);
            acc[k & 3] = (acc[k & 3] | x) ^ a[(k + 7) & 31];   // one 3-input LOP3
            a[k] = x;                                           // every value changes every iteration
        }
    }
    long long t1 = clock64();
    uint32_t r = acc[0] ^ acc[1] ^ acc[2] ^ acc[3];
    for (int k = 0; k < 32; k++) r ^= a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (unsigned long long)(t1 - t0);
}

int main() {
    uint32_t* out; unsigned long long* cyc;
    cudaMalloc(&out, 148 * 4 * 512 * 4); cudaMallocManaged(&cyc, 8);
    const int iters = 20000;
    const char* names[5] = {"32 SHF + 32 LOP3", "32 IMAD.HI/IMAD pairs + 32 LOP3", "16 SHF + 16 pairs + 32 LOP3", "64 LOP3", "24 SHF + 8 pairs + 32 LOP3"};
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int mode = 0; mode < 5; mode++) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) bench<0><<<148, warps * 32>>>(out, iters, 13, cyc);
                if (mode == 1) bench<1><<<148, warps * 32>>>(out, iters, 13, cyc);
                if (mode == 2) bench<2><<<148, warps * 32>>>(out, iters, 13, cyc);
                if (mode == 3) bench<3><<<148, warps * 32>>>(out, iters, 13, cyc);
                if (mode == 4) bench<4><<<148, warps * 32>>>(out, iters, 13, cyc);
                cudaDeviceSynchronize();
            }
            printf("warps/SM %2d  %-34s %7.1f clk per iteration -> %6.1f clk per warp-iteration per SMSP\n",
                   warps, names[mode], (double)*cyc / iters, (double)*cyc / iters / (warps / 4.0));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
