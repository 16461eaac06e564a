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
//     i.e. one funnel shift + one LOP3 per plane: 16 integer instructions for
//     32 candidate-compares, where a thread-per-position kernel needs 32 loads
//     and 32 compares.
//   * "A match of >= 3 starts at i" is E & E>>1 & E>>2 (bits shifted in from the
//     next block); >= 5 and >= 9 follow by doubling.  Per-position state is kept
//     as bit masks too: c5/c9 = "this position already holds a match of >= 4 / >= 8
//     and needs >= 5 / >= 9 to improve", dn = "holds max_len, finished".
//   * Distances are visited in ascending order, exactly like the reference, so
//     "strictly longer wins" keeps the nearest candidate among equals.
//   * Only positions that survive the mask test (a few per position over the
//     whole 32767-distance scan) reach the scalar path, which
//     measures the run exactly from the same E bits, compares it with the
//     position's current best and records (len, dist).
//
// Work split: a thread owns Q consecutive blocks (32*Q positions) for the whole
// scan, so all per-position state is private to one thread: no atomics, no
// inter-thread ordering.  A warp covers 32*Q blocks of which the last one is a
// halo (needed only as look-ahead; it is the next warp's first block).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace v2 {

constexpr int kWarps = 4;                 // warps per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kQ = 4;                     // blocks of 32 positions per thread
constexpr int kWarpBlocks = 32 * kQ;      // blocks a warp computes (last one is halo)
constexpr int kWarpOwned = kWarpBlocks - 1;
constexpr int kTileBlocks = kWarps * kWarpOwned;      // owned blocks per CTA
constexpr int kTilePos = kTileBlocks * 32;            // positions per CTA

struct Geometry {            // host-computed, identical for all CTAs of a launch
    int back_blocks;         // plane blocks staged before the tile: ceil(max_dist/32) + 1
    int ahead_blocks;        // plane blocks staged after the tile + halo (long-run extension)
    int plane_blocks;        // back_blocks + kTileBlocks + 1 + ahead_blocks
    int smem_bytes;
};

__host__ __device__ inline Geometry geometry(uint32_t max_len, uint32_t max_dist, bool edge) {
    Geometry g;
    g.back_blocks = (int)((max_dist + 31) / 32) + 1;
    g.ahead_blocks = (int)((max_len + 31) / 32) + 2;
    g.plane_blocks = g.back_blocks + kTileBlocks + 1 + g.ahead_blocks;
    int bytes = g.plane_blocks * 32;                  // 8 planes x 4 B per block
    bytes += (kTilePos + 32) * 2;                     // u16 best length per owned position
    if (edge) { bytes += g.plane_blocks * 4; }        // validity plane
    g.smem_bytes = (bytes + 15) & ~15;
    return g;
}

__device__ __forceinline__ uint32_t fsr(uint32_t lo, uint32_t hi, int s) {
    return __funnelshift_r(lo, hi, s);               // bits [s, s+32) of hi:lo
}

// E-bar (1 = bytes differ) of plane block `blk` (smem block index) at distance
// 32*m - sh, straight from shared memory.  Scalar path only.
template <bool kEdge>
__device__ __forceinline__ uint32_t ebar_from_smem(const uint4* __restrict__ PL,
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

// kMinLen: 2 or 3.  Levels: L0 = kMinLen, L1 = L0 + kS1, L2 = L1 + kS2.
template <int kMinLen, bool kEdge>
__global__ void __launch_bounds__(kThreads)
match_table(const uint8_t* __restrict__ shard, long long back, long long n, long long ahead,
            uint32_t max_len, uint32_t max_dist, uint32_t* __restrict__ table,
            const int* __restrict__ tile_list, int tile_first) {
    constexpr int kS1 = kMinLen - 1;                  // 3 -> 5, 2 -> 3
    constexpr int kL1 = kMinLen + kS1;
    constexpr int kS2 = kL1 - 1;                      // 5 -> 9, 3 -> 5
    constexpr int kL2 = kL1 + kS2;

    extern __shared__ __align__(16) uint8_t smem_raw[];
    const Geometry geo = geometry(max_len, max_dist, kEdge);
    uint4* PL = reinterpret_cast<uint4*>(smem_raw);                       // [plane_blocks][2]
    uint16_t* best_len = reinterpret_cast<uint16_t*>(smem_raw + geo.plane_blocks * 32);
    uint32_t* VL = reinterpret_cast<uint32_t*>(smem_raw + geo.plane_blocks * 32 + (kTilePos + 32) * 2);

    const int tile = tile_list != nullptr ? tile_list[blockIdx.x] : tile_first + (int)blockIdx.x;
    const long long tile_pos0 = (long long)tile * kTilePos;              // shard-relative
    const long long plane_pos0 = tile_pos0 - (long long)geo.back_blocks * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- stage: bytes -> bit planes (and validity), zero the outputs ---------
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
        for (int k = threadIdx.x; k < kTilePos + 32; k += kThreads) { best_len[k] = 0; }
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
    uint32_t c1[kQ], c2[kQ], dn[kQ];
#pragma unroll
    for (int q = 0; q < kQ; q++) {
        const uint4 a = PL[2 * (blk0 + q)], b = PL[2 * (blk0 + q) + 1];
        qv[q][0] = a.x; qv[q][1] = a.y; qv[q][2] = a.z; qv[q][3] = a.w;
        qv[q][4] = b.x; qv[q][5] = b.y; qv[q][6] = b.z; qv[q][7] = b.w;
        c1[q] = 0; c2[q] = 0;
        // closed from the start: the warp's halo block and positions past the shard
        const long long p0 = tile_pos0 + (long long)(own0 + q) * 32;
        uint32_t closed = 0;
        if (lane == 31 && q == kQ - 1) { closed = 0xFFFFFFFFu; }
        else if (p0 + 32 > n) { closed = p0 >= n ? 0xFFFFFFFFu : (0xFFFFFFFFu << (int)(n - p0)); }
        dn[q] = closed;
        vq[q] = kEdge ? VL[blk0 + q] : 0xFFFFFFFFu;
    }

    // farthest distance any position of this tile can use
    const long long tile_last = min(tile_pos0 + kTilePos + 31, n - 1);
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
#pragma unroll 2
        for (int sh = 31; sh >= 0; sh--) {
            const uint32_t d = (uint32_t)(32 * m - sh);
            if (d > reach) { break; }
            uint32_t eb[kQ + 1], r0[kQ + 1], r1[kQ + 1], r2[kQ];
#pragma unroll
            for (int q = 0; q < kQ; q++) {
                uint32_t e = fsr(cw[q][0], cw[q + 1][0], sh) ^ qv[q][0];
#pragma unroll
                for (int b = 1; b < 8; b++) { e |= fsr(cw[q][b], cw[q + 1][b], sh) ^ qv[q][b]; }
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
#pragma unroll
            for (int q = 0; q < kQ; q++) { r1[q] = r0[q] | fsr(r0[q], r0[q + 1], kS1); }
            r1[kQ] = __shfl_down_sync(0xFFFFFFFFu, r1[0], 1);
            uint32_t none = 0xFFFFFFFFu;
            uint32_t ib[kQ];
#pragma unroll
            for (int q = 0; q < kQ; q++) {
                r2[q] = r1[q] | fsr(r1[q], r1[q + 1], kS2);
                ib[q] = r0[q] | (r1[q] & c1[q]) | (r2[q] & c2[q]) | dn[q];
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
                        const int k = (own0 + q) * 32 + p;              // tile-relative position
                        const uint32_t have = best_len[k];
                        const uint32_t win = fsr(eb[q], eb[q + 1], p);  // E-bar from position p on
                        uint32_t run;
                        if (win != 0) {
                            run = (uint32_t)(__ffs((int)win) - 1);      // run ends inside the window
                            if (run <= have) { continue; }
                        } else {
                            // the run leaves the 32-bit window: keep measuring from the planes
                            run = 32;
                            int blk = blk0 + q + 1;
                            while (run < max_len) {
                                const uint32_t ea = ebar_from_smem<kEdge>(PL, VL, blk, m, sh);
                                const uint32_t ec = ebar_from_smem<kEdge>(PL, VL, blk + 1, m, sh);
                                const uint32_t w2 = fsr(ea, ec, p);
                                if (w2 != 0) { run += (uint32_t)(__ffs((int)w2) - 1); break; }
                                run += 32;
                                blk++;
                            }
                        }
                        run = min(run, max_len);
                        if (run <= have) { continue; }
                        best_len[k] = (uint16_t)run;
                        table[tile_pos0 + k] = (run << 16) | d;
                        const uint32_t bit = 1u << p;
                        if (run >= (uint32_t)(kL1 - 1)) { c1[q] |= bit; }
                        if (run >= (uint32_t)(kL2 - 1)) { c2[q] |= bit; }
                        if (run >= max_len) { dn[q] |= bit; }
                    }
                }
            }
        }
    }
}

}  // namespace v2
