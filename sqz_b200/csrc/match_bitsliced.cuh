// match_bitsliced.cuh -- the tuned match-table kernel of sqz-b200 (sm_100a).
//
// Same result as the reference's brute-force scan
// (/root/reference/attic/map_experiment/squeeze.h:338-358 at every position),
// computed 32 positions at a time with bit-parallel logic:
//
//   * The input is transposed into 8 bit-planes in shared memory: bit i of
//     plane b is bit b of byte i.  For a block of 32 consecutive positions and a
//     distance d, "byte i equals byte i-d" for all 32 positions is
//         E = AND_b ~( plane_b[i..i+31] ^ plane_b[i-d..i-d+31] )
//     i.e. one funnel shift + one LOP3 per plane: 17 integer instructions for
//     32 candidate-compares, where a thread-per-position kernel needs 32 loads
//     and 32 compares.
//   * "A match of >= 3 starts at i" is E & E>>1 & E>>2 (bits shifted in from the
//     next block); >= 5, 9, 17, 33, 65 follow by doubling (R_2k-1 = R_k & R_k>>(k-1)).
//   * Per-position state is bit-sliced too: three code planes hold, for each of
//     the 32 positions, which of the six run-length classes a candidate has to
//     reach to beat the position's current best (codes 6/7 = closed: the
//     position already holds max_len, or is not owned by this thread).  A 6-way
//     bit-wise multiplexer picks the matching run mask for every position.
//   * Distances are visited in ascending order, exactly like the reference, so
//     "strictly longer wins" keeps the nearest candidate among equals.
//   * Only positions that survive the multiplexer reach the scalar path, which
//     measures the run exactly from the same E bits, compares it with the
//     position's current best (kept in the output table itself) and records
//     (len, dist).
//
// Work split: a thread owns kQ consecutive blocks (32*kQ positions) for the
// whole scan, so all per-position state is private to one thread: no atomics,
// no inter-thread ordering.  Lane 31 of every warp recomputes the first kQ
// blocks of the next warp as look-ahead only (its positions are closed).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace v2 {

constexpr int kWarps = 4;                 // warps per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kQ = 4;                     // blocks of 32 positions per thread
constexpr int kWarpOwned = 31 * kQ;       // blocks a warp owns (lane 31 is look-ahead)
constexpr int kTileBlocks = kWarps * kWarpOwned;      // owned blocks per CTA
constexpr int kTilePos = kTileBlocks * 32;            // positions per CTA
constexpr int kLevels = 6;

struct Geometry {            // identical for all CTAs of a launch
    int back_blocks;         // plane blocks staged before the tile: ceil(max_dist/32) + 1
    int ahead_blocks;        // after the tile: look-ahead lane + long-run extension
    int plane_blocks;
    int smem_bytes;
};

__host__ __device__ inline Geometry geometry(uint32_t max_len, uint32_t max_dist, bool edge) {
    Geometry g;
    g.back_blocks = (int)((max_dist + 31) / 32) + 1;
    g.ahead_blocks = kQ + (int)((max_len + 31) / 32) + 3;
    g.plane_blocks = g.back_blocks + kTileBlocks + g.ahead_blocks;
    int bytes = g.plane_blocks * 32;                  // 8 planes x 4 B per block
    if (edge) { bytes += g.plane_blocks * 4; }        // validity plane
    g.smem_bytes = (bytes + 15) & ~15;
    return g;
}

__device__ __forceinline__ uint32_t fsr(uint32_t lo, uint32_t hi, int s) {
    return __funnelshift_r(lo, hi, s);               // bits [s, s+32) of hi:lo
}

// bit-wise 2:1 multiplexer: sel ? b : a
__device__ __forceinline__ uint32_t mux(uint32_t sel, uint32_t a, uint32_t b) {
    return (a & ~sel) | (b & sel);
}

// E-bar (1 = bytes differ) of plane block `blk` at distance 32*m - sh, straight
// from shared memory.  Scalar path only.
template <bool kEdge>
__device__ __noinline__ uint32_t ebar_from_smem(const uint4* __restrict__ PL,
                                                const uint32_t* __restrict__ VL,
                                                int blk, int m, int sh) {
    const uint4 qa = PL[2 * blk], qb = PL[2 * blk + 1];
    const uint4 la = PL[2 * (blk - m)], lb = PL[2 * (blk - m) + 1];
    const uint4 ha = PL[2 * (blk - m + 1)], hb = PL[2 * (blk - m + 1) + 1];
    uint32_t e = fsr(la.x, ha.x, sh) ^ qa.x;
    e |= fsr(la.y, ha.y, sh) ^ qa.y;
    e |= fsr(la.z, ha.z, sh) ^ qa.z;
    e |= fsr(la.w, ha.w, sh) ^ qa.w;
    e |= fsr(lb.x, hb.x, sh) ^ qb.x;
    e |= fsr(lb.y, hb.y, sh) ^ qb.y;
    e |= fsr(lb.z, hb.z, sh) ^ qb.z;
    e |= fsr(lb.w, hb.w, sh) ^ qb.w;
    if (kEdge) { e |= ~VL[blk] | ~fsr(VL[blk - m], VL[blk - m + 1], sh); }
    return e;
}

// class a position is in once it holds a match of length `best`:
// thresholds kMinLen, then (T-1)*2+1 each: 3,5,9,17,33,65 (or 2,3,5,9,17,33)
template <int kMinLen>
__device__ __forceinline__ int level_of(uint32_t best) {
    if (best < 2) { return 0; }
    const int lg = 31 - __clz((int)best);
    return min(kLevels - 1, max(0, lg - (kMinLen - 2)));
}

template <int kMinLen, bool kEdge>
__global__ void __launch_bounds__(kThreads)
match_table(const uint8_t* __restrict__ shard, long long back, long long n, long long ahead,
            uint32_t max_len, uint32_t max_dist, uint32_t* __restrict__ table, int tile_first,
            unsigned long long* __restrict__ tile_cycles) {
    const long long t_begin = clock64();
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const Geometry geo = geometry(max_len, max_dist, kEdge);
    const uint4* PL = reinterpret_cast<const uint4*>(smem_raw);           // [plane_blocks][2]
    uint32_t* VL = reinterpret_cast<uint32_t*>(smem_raw + geo.plane_blocks * 32);

    const int tile = tile_first + (int)blockIdx.x;
    const long long tile_pos0 = (long long)tile * kTilePos;              // shard-relative
    const long long plane_pos0 = tile_pos0 - (long long)geo.back_blocks * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- stage: bytes -> bit planes (and validity), zero this tile's outputs ----
    {
        uint32_t* PLw = reinterpret_cast<uint32_t*>(smem_raw);
        for (int blk = warp; blk < geo.plane_blocks; blk += kWarps) {
            const long long pos = plane_pos0 + (long long)blk * 32 + lane;
            const bool ok = pos >= -back && pos < n + ahead;
            const uint32_t byte = ok ? (uint32_t)__ldg(shard + pos) : 0u;
            uint32_t mine = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const uint32_t w = __ballot_sync(0xFFFFFFFFu, (byte >> b) & 1u);
                if (lane == b) { mine = w; }
            }
            if (lane < 8) { PLw[blk * 8 + lane] = mine; }
            if (kEdge) {
                const uint32_t v = __ballot_sync(0xFFFFFFFFu, ok);
                if (lane == 0) { VL[blk] = v; }
            }
        }
        for (int k = threadIdx.x; k < kTilePos; k += kThreads) {
            const long long p = tile_pos0 + k;
            if (p < n) { table[p] = 0; }
        }
    }
    __syncthreads();

    // ---- per-thread state ------------------------------------------------------
    const int own0 = warp * kWarpOwned + lane * kQ;         // first block of this thread (tile-relative)
    const int blk0 = geo.back_blocks + own0;                // same, as plane block index
    uint32_t qv[kQ][8];
    uint32_t vq[kQ];
    uint32_t L0[kQ], L1[kQ], L2[kQ];                        // bit-sliced class code per position
#pragma unroll
    for (int q = 0; q < kQ; q++) {
        const uint4 a = PL[2 * (blk0 + q)], b = PL[2 * (blk0 + q) + 1];
        qv[q][0] = a.x; qv[q][1] = a.y; qv[q][2] = a.z; qv[q][3] = a.w;
        qv[q][4] = b.x; qv[q][5] = b.y; qv[q][6] = b.z; qv[q][7] = b.w;
        // closed from the start: the look-ahead lane and positions past the shard
        const long long p0 = tile_pos0 + (long long)(own0 + q) * 32;
        uint32_t closed = 0;
        if (lane == 31) { closed = 0xFFFFFFFFu; }
        else if (p0 + 32 > n) { closed = p0 >= n ? 0xFFFFFFFFu : (0xFFFFFFFFu << (int)(n - p0)); }
        L0[q] = 0; L1[q] = closed; L2[q] = closed;          // code 6 = closed
        vq[q] = kEdge ? VL[blk0 + q] : 0xFFFFFFFFu;
    }

    // farthest distance any position of this tile can use
    const long long tile_last = min(tile_pos0 + kTilePos - 1, n - 1);
    const uint32_t reach = (uint32_t)min((long long)max_dist, tile_last + back);
    const int m_end = (int)((reach + 31) / 32);

    for (int m = 1; m <= m_end; m++) {
        uint32_t cw[kQ + 1][8];
        uint32_t vc[kQ + 1];
#pragma unroll
        for (int j = 0; j <= kQ; j++) {
            const uint4 a = PL[2 * (blk0 - m + j)], b = PL[2 * (blk0 - m + j) + 1];
            cw[j][0] = a.x; cw[j][1] = a.y; cw[j][2] = a.z; cw[j][3] = a.w;
            cw[j][4] = b.x; cw[j][5] = b.y; cw[j][6] = b.z; cw[j][7] = b.w;
            vc[j] = kEdge ? VL[blk0 - m + j] : 0xFFFFFFFFu;
        }
#pragma unroll 1
        for (int sh = 31; sh >= 0; sh--) {
            const uint32_t d = (uint32_t)(32 * m - sh);
            if (d > reach) { break; }
            // r[k][q]: bit p set = NO run of at least T_k equal bytes starts at position p
            uint32_t eb[kQ + 1], r0[kQ + 1], r1[kQ + 1], r2[kQ + 1], r3[kQ + 1], r4[kQ + 1];
#pragma unroll
            for (int q = 0; q < kQ; q++) {
                uint32_t ea = fsr(cw[q][0], cw[q + 1][0], sh) ^ qv[q][0];
                uint32_t ec = fsr(cw[q][4], cw[q + 1][4], sh) ^ qv[q][4];
#pragma unroll
                for (int b = 1; b < 4; b++) {
                    ea |= fsr(cw[q][b], cw[q + 1][b], sh) ^ qv[q][b];
                    ec |= fsr(cw[q][b + 4], cw[q + 1][b + 4], sh) ^ qv[q][b + 4];
                }
                uint32_t e = ea | ec;
                if (kEdge) { e |= ~vq[q] | ~fsr(vc[q], vc[q + 1], sh); }
                eb[q] = e;
            }
            eb[kQ] = __shfl_down_sync(0xFFFFFFFFu, eb[0], 1);
#pragma unroll
            for (int q = 0; q < kQ; q++) {
                uint32_t r = eb[q] | fsr(eb[q], eb[q + 1], 1);
                if (kMinLen >= 3) { r |= fsr(eb[q], eb[q + 1], 2); }
                r0[q] = r;
            }
            r0[kQ] = __shfl_down_sync(0xFFFFFFFFu, r0[0], 1);
            constexpr int S1 = kMinLen - 1, S2 = 2 * S1, S3 = 2 * S2, S4 = 2 * S3, S5 = 2 * S4;
#pragma unroll
            for (int q = 0; q < kQ; q++) { r1[q] = r0[q] | fsr(r0[q], r0[q + 1], S1); }
            r1[kQ] = __shfl_down_sync(0xFFFFFFFFu, r1[0], 1);
#pragma unroll
            for (int q = 0; q < kQ; q++) { r2[q] = r1[q] | fsr(r1[q], r1[q + 1], S2); }
            r2[kQ] = __shfl_down_sync(0xFFFFFFFFu, r2[0], 1);
#pragma unroll
            for (int q = 0; q < kQ; q++) { r3[q] = r2[q] | fsr(r2[q], r2[q + 1], S3); }
            r3[kQ] = __shfl_down_sync(0xFFFFFFFFu, r3[0], 1);
#pragma unroll
            for (int q = 0; q < kQ; q++) { r4[q] = r3[q] | fsr(r3[q], r3[q + 1], S4); }
            r4[kQ] = __shfl_down_sync(0xFFFFFFFFu, r4[0], 1);
            uint32_t ib[kQ];
            uint32_t none = 0xFFFFFFFFu;
#pragma unroll
            for (int q = 0; q < kQ; q++) {
                const uint32_t r5 = r4[q] | (S5 >= 32 ? r4[q + 1] : fsr(r4[q], r4[q + 1], S5 & 31));
                const uint32_t m01 = mux(L0[q], r0[q], r1[q]);
                const uint32_t m23 = mux(L0[q], r2[q], r3[q]);
                const uint32_t m45 = mux(L0[q], r4[q], r5) | L1[q];      // codes 6,7: closed
                const uint32_t m03 = mux(L1[q], m01, m23);
                ib[q] = mux(L2[q], m03, m45);
                none &= ib[q];
            }
            if (none != 0xFFFFFFFFu) {
                // ---- scalar path: exact decision for the few surviving positions ----
#pragma unroll
                for (int q = 0; q < kQ; q++) {
                    uint32_t todo = ~ib[q];
                    while (todo != 0) {
                        const int p = __ffs((int)todo) - 1;
                        todo &= todo - 1;
                        uint32_t* slot = table + tile_pos0 + (long long)(own0 + q) * 32 + p;
                        const uint32_t have = *slot >> 16;
                        const uint32_t win = fsr(eb[q], eb[q + 1], p);  // E-bar from position p on
                        uint32_t run;
                        if (win != 0) {
                            run = (uint32_t)(__ffs((int)win) - 1);      // run ends inside the window
                            if (run <= have) { continue; }
                        } else {
                            // the run leaves the 32-bit window
                            if (have >= 32) {
                                // cheap reject first: the byte at offset `have` has to match too
                                const int off = p + (int)have;
                                const uint32_t e = ebar_from_smem<kEdge>(PL, VL, blk0 + q + (off >> 5), m, sh);
                                if ((e >> (off & 31)) & 1u) { continue; }
                            }
                            run = 32;
                            int blk = blk0 + q + 1;
                            uint32_t ea = ebar_from_smem<kEdge>(PL, VL, blk, m, sh);
                            while (run < max_len) {
                                const uint32_t ec = ebar_from_smem<kEdge>(PL, VL, blk + 1, m, sh);
                                const uint32_t w2 = fsr(ea, ec, p);
                                if (w2 != 0) { run += (uint32_t)(__ffs((int)w2) - 1); break; }
                                run += 32;
                                blk++;
                                ea = ec;
                            }
                        }
                        run = min(run, max_len);
                        if (run <= have) { continue; }
                        *slot = (run << 16) | d;
                        const uint32_t bit = 1u << p;
                        const int code = run >= max_len ? 6 : level_of<kMinLen>(run);
                        L0[q] = (code & 1) ? (L0[q] | bit) : (L0[q] & ~bit);
                        L1[q] = (code & 2) ? (L1[q] | bit) : (L1[q] & ~bit);
                        L2[q] = (code & 4) ? (L2[q] | bit) : (L2[q] & ~bit);
                    }
                }
            }
        }
    }
    if (tile_cycles != nullptr) {          // debugging aid: per-tile duration
        __syncthreads();
        if (threadIdx.x == 0) { tile_cycles[tile] = (unsigned long long)(clock64() - t_begin); }
    }
}

}  // namespace v2
