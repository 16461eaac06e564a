// match_bitsliced.cuh -- the tuned match-table kernel of sqz-b200 (sm_100a).
//
// Same result as the reference's brute-force scan
// (/root/reference/attic/map_experiment/squeeze.h:338-358 at every position),
// computed 32 positions at a time with bit-parallel logic.
//
// Phase 1 (all 32767 distances, every position, bit-sliced):
//   * The input is transposed into 8 bit-planes in shared memory: bit i of
//     plane b is bit b of byte i.  For a block of 32 consecutive positions and a
//     distance d, "byte i equals byte i-d" for all 32 positions is
//         E = AND_b ~( plane_b[i..i+31] ^ plane_b[i-d..i-d+31] )
//     i.e. one funnel shift + one LOP3 per plane for 32 candidate-compares, where
//     a thread-per-position kernel needs 32 loads and 32 compares.
//   * Per-position state is bit-sliced too: the run length a candidate has to
//     reach to beat the position's current best, need = best+1 (exact up to
//     min_len+3, anything longer counts as min_len+3), is kept as a thermometer
//     of three masks G_k = "byte offset min_len+k has to match as well".  A
//     candidate fails at position p iff one of its first need(p) bytes differs:
//         fail = E' | E'>>1 | E'>>2 | (E'>>3 & G_0) | (E'>>4 & G_1) | (E'>>5 & G_2) | closed
//     (E' = ~E, bits shifted in from the next block), so a candidate that merely
//     ties or falls short never leaves the fast path.
//   * Distances ascend group by group (128 at a time) like the reference's scan; the
//     four distances 32 apart inside a group share their shifted candidate words
//     (7 funnel shifts per plane for 16 block-distance pairs).  Positions whose
//     best was set inside the current group accept a nearer equal run, so the
//     result is still "longest, nearest among equals".
//   * Survivors of that test take a scalar path: the run is measured from the
//     same E bits (32-bit window), compared with the position's current best and
//     recorded as (len, dist).
//
// Phase 2 (few positions, separate kernel finish_marked):
//   Positions whose run leaves the 32-bit window (matches of >= 32 bytes, the
//   ends of long byte runs) or that keep producing near-ties are marked in the
//   table and closed afterwards: by inheritance from the position above where
//   that is provably exact (every position inside a long match but its last),
//   else by the classic exact search with one warp per position -- 32 lanes x 4
//   candidates per step, 4-byte compares at the offset a candidate must match to
//   win and at the most distinctive window of the bytes already matched, ballot
//   for the nearest hit, cooperative verify.
//
// Work split in phase 1: a thread owns kQ consecutive blocks (32*kQ positions)
// for the whole scan, so all per-position state is private to one thread: no
// atomics, no inter-thread ordering.  The last block of every warp is the first
// block of the next warp, recomputed as look-ahead only (its positions are closed).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

// Bounds-checked build (-DSQZ_BOUNDS, tools/sanitize_case.py): every indexed access of the
// kernels is range-checked and traps with a message.  compute-sanitizer is not available on
// the GPU pool, so this is how out-of-range accesses are looked for.
#ifdef SQZ_BOUNDS
#include <cstdio>
#define SQZ_CHECK(cond, what)                                                             \
    do {                                                                                  \
        if (!(cond)) { printf("SQZ_BOUNDS: %s (line %d)\n", what, __LINE__); __trap(); }  \
    } while (0)
#else
#define SQZ_CHECK(cond, what) do { } while (0)
#endif

namespace v2 {

constexpr int kWarps = 4;                 // warps per CTA
constexpr int kThreads = kWarps * 32;
// Blocks of 32 positions per thread (kQ, a template parameter of the phase 1 kernel).  4 is the
// throughput shape: four distances share their shifted words, 31 candidate-compares per instruction.
// 1 is the latency shape for small shards: a quarter of the positions and of the scalar-path work
// per warp, four times as many warps, at 1.4 times the instructions per candidate-compare.
__host__ __device__ constexpr int warp_owned(int q) { return 32 * q - 1; }          // blocks a warp owns; its last block is look-ahead only
__host__ __device__ constexpr int tile_blocks(int q) { return kWarps * warp_owned(q); }   // owned blocks per CTA
__host__ __device__ constexpr int tile_pos(int q) { return tile_blocks(q) * 32; }         // positions per CTA
// staging piece: its bytes + alignment slack fit in the state bytes of the tile
__host__ __device__ constexpr int stage_blocks(int q) { return tile_blocks(q) - 4 < 448 ? tile_blocks(q) - 4 : 448; }
#ifndef SQZ_GATED
#define SQZ_GATED 3
#endif
#ifndef SQZ_GATED_Q1
#define SQZ_GATED_Q1 3
#endif
// need is tracked exactly up to min_len + gated(q): three levels measure best for the throughput
// shape (each level costs 4-5 % of its loop)
__host__ __device__ constexpr int gated(int q) { return q == 1 ? SQZ_GATED_Q1 : SQZ_GATED; }
// Bit planes 0..kQuietPlanes-1 hold a linear hash of each byte -- its low bits XOR a constant chosen by
// its high bits -- and the remaining planes the high bits as they are: hash and high bits together
// determine the byte, so a compare over all eight planes is the byte compare it always was.  A compare
// over the hashed planes alone lets every candidate through that the exact one would, plus those whose
// bytes differ by one of three patterns: 0.12 candidates per position over the whole window on random
// bytes, 5-7 on ELF (contexts like 00 00 xx carry eight bits, six after hashing;
// tools/plane_hash_false_positives.c).  A warp whose data is quiet -- the same warps that drop the need masks,
// see the adaptive gate -- compares the hashed planes only: 46 of the 246 ALU instructions of its loop
// body.  The scalar path measures runs on all eight planes, so every decision stays exact.
#ifndef SQZ_QUIET_PLANES
#define SQZ_QUIET_PLANES 6
#endif
constexpr int kQuietPlanes = SQZ_QUIET_PLANES;
static_assert(kQuietPlanes >= 6 && kQuietPlanes <= 8, "6, 7 or 8 planes");
// byte k of the constant = what is XORed into a byte whose high 8-kQuietPlanes bits are k
constexpr uint32_t kFold = kQuietPlanes == 6 ? (0x2Bu << 8 | 0x16u << 16 | (0x2Bu ^ 0x16u) << 24) : kQuietPlanes == 7 ? (0x1Du << 8) : 0u;
__host__ __device__ constexpr uint32_t fold_byte(uint32_t byte) { return byte ^ ((kFold >> (8 * (byte >> kQuietPlanes))) & 0xFFu); }

#ifndef SQZ_BUSY_ITERATIONS
#define SQZ_BUSY_ITERATIONS 8
#endif
#ifndef SQZ_QUIET_ENTRIES
#define SQZ_QUIET_ENTRIES 5
#endif
#ifndef SQZ_BUSY_ENTRIES
#define SQZ_BUSY_ENTRIES 10
#endif
#ifndef SQZ_T0_STRICT
#define SQZ_T0_STRICT 1   // a best found at the nearest of an iteration's distances (t = 0) is strict at once
#endif
#ifndef SQZ_TIE_MASK
#define SQZ_TIE_MASK 0
#endif
constexpr uint32_t kTieMask = SQZ_TIE_MASK;   // a rejected survivor is counted when (d & kTieMask) == 0; the 7th hands the position over
#ifndef SQZ_TIE_LIMIT
#define SQZ_TIE_LIMIT 6
#endif
constexpr uint32_t kTieLimit = SQZ_TIE_LIMIT; // rejected survivors a position may collect before the next one hands it over (<= 6)
#ifndef SQZ_SLICED_TIE_LIMIT
#define SQZ_SLICED_TIE_LIMIT 6
#endif
constexpr uint32_t kSlicedTieLimit = SQZ_SLICED_TIE_LIMIT;   // the same in a sliced launch (8 = never hand over on rejects)
constexpr uint8_t kHandOver = 0xFF;       // best_len mark: finish this position in phase 2
constexpr uint32_t kOpenBit = 0x80000000u; // table word mark: position is finished by the phase 2 kernel
constexpr int kResumeShift = 26;          // bits 26..30 of an open word: distance/1024 below which all is settled
constexpr uint32_t kStateMask = 0x03FFFFFFu; // (len << 16) | dist of an open word

struct Geometry {            // identical for all CTAs of a launch
    int back_blocks;         // plane blocks staged before the tile: ceil(max_dist/32) + 2*kQ
    int ahead_blocks;        // after the tile: look-ahead block + slack
    int plane_blocks;
    int region_bytes;        // the bit planes
    int smem_bytes;
};

__host__ __device__ inline Geometry geometry(uint32_t max_len, uint32_t max_dist, bool edge, bool seeded, int q) {
    const int kTileBlocks = tile_blocks(q), kTilePos = tile_pos(q);
    Geometry g;
    g.back_blocks = (int)((max_dist + 31) / 32) + 2 * q;    // a group of q word distances + its window
    g.ahead_blocks = q + 2;
    g.plane_blocks = g.back_blocks + kTileBlocks + g.ahead_blocks;
    (void)max_len;
    g.region_bytes = g.plane_blocks * 32;                                  // 8 planes x 4 B per block
    int bytes = g.region_bytes + kTilePos + 32;                            // + one state byte per position
    if (edge) { bytes += g.plane_blocks * 4; }                             // validity plane
    if (seeded) { bytes += (kTileBlocks + 1) * (gated(q) + 1) * 4; }       // starting masks of a seeded slice
    g.smem_bytes = (bytes + 15) & ~15;
    return g;
}

__device__ __forceinline__ uint32_t fsr(uint32_t lo, uint32_t hi, int s) {
    return __funnelshift_r(lo, hi, s);               // bits [s, s+32) of hi:lo
}

// a | (b ^ c) as the one LOP3 it is (left to itself the compiler regroups the two planes the gated body
// adds into three instructions per pair instead of two)
__device__ __forceinline__ uint32_t or_xor(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xF6;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Planes kQuietPlanes..7 (the bytes' high bits as they are) of the pair (query block qb, candidate blocks cb, cb+1
// shifted by sh): what a compare over the hashed planes cannot tell.  Shared memory; the scalar path only.
__device__ __forceinline__ uint32_t high_planes_differ(const uint4* __restrict__ PL, int qb, int cb, int sh) {
    const uint32_t* W = reinterpret_cast<const uint32_t*>(PL);
    uint32_t x = 0;
#pragma unroll
    for (int b = kQuietPlanes; b < 8; b++) { x |= fsr(W[8 * cb + b], W[8 * (cb + 1) + b], sh) ^ W[8 * qb + b]; }
    return x;
}

// the eight plane words of plane block `blk` (shared memory)
__device__ __forceinline__ void load_planes(const uint4* __restrict__ PL, int blk, uint32_t (&w)[8]) {
    const uint4 a = PL[2 * blk], b = PL[2 * blk + 1];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}

// ---------------------------------------------------------------------------
// phase 2: one warp finishes one position with the exact search on raw bytes.
// S = 4-byte aligned byte image, S[xi] = the position's first
// byte, x_end = one past the last readable byte; candidates S[xi-d] for d up to
// `reach`; `room` = min(max_len, bytes left).  (best, bdist) enter with what
// phase 1 found (all distances < max(bdist+1, d_from) are settled) and leave final.
//
// 32 lanes x 2 words x 4 candidates per step: every lane takes two aligned words and tests
// the four byte offsets in each against the 4 bytes a candidate has to match to
// win (the word ending at offset need-1; the first `need` bytes while
// need < 4).  Hits are verified nearest first by all lanes together.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t word_at(const uint32_t* __restrict__ W, int w, int w_last) {
    return w <= w_last ? W[w] : 0u;
}

// 4 bytes starting at byte index x of the image
__device__ __forceinline__ uint32_t bytes_at(const uint32_t* __restrict__ W, int x, int w_last) {
    const int w = x >> 2;
    return fsr(word_at(W, w, w_last), word_at(W, w + 1, w_last), (x & 3) * 8);
}

// Among the 4-byte windows inside the first `best` bytes of the position pick the one a
// random candidate is least likely to share: most distinct byte values, earliest on ties
// (far from the window at the end that the main filter already tests).  -1 if none fits.
__device__ __forceinline__ int pick_second_window(const uint32_t* __restrict__ W, int xi, uint32_t best,
                                                  int w_last, int lane) {
    if (best < 4) { return -1; }
    uint32_t mine = 0;
    for (int k = lane; k + 4 <= (int)best; k += 32) {
        const uint32_t x = bytes_at(W, xi + k, w_last);
        const uint32_t b0 = x & 0xFF, b1 = (x >> 8) & 0xFF, b2 = (x >> 16) & 0xFF, b3 = x >> 24;
        const uint32_t score = (b0 != b1) + (b0 != b2) + (b0 != b3) + (b1 != b2) + (b1 != b3) + (b2 != b3);
        mine = max(mine, ((score + 1) << 16) | (uint32_t)(0xFFFF - k));
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) { mine = max(mine, __shfl_xor_sync(0xFFFFFFFFu, mine, sft)); }
    return 0xFFFF - (int)(mine & 0xFFFFu);
}

__device__ __forceinline__ void finish_position(const uint8_t* __restrict__ S, int xi, int x_end,
                                             uint32_t d_from, uint32_t reach, uint32_t room, uint32_t min_len,
                                             uint32_t& best, uint32_t& bdist, int lane,
                                             unsigned long long* dbg = nullptr, uint32_t* runner_up = nullptr) {
    const uint32_t* W = reinterpret_cast<const uint32_t*>(S);
    const int w_last = (x_end - 1) >> 2;
    unsigned int n_steps = 0, n_verify = 0, n_rounds = 0, n_improve = 0;
    uint32_t d0 = max(bdist + 1, d_from);              // everything nearer is settled
    while (d0 <= reach && best < room) {
        const uint32_t need = max(best + 1, min_len);
        if (need > room) { break; }
        // main filter: the 4 bytes ending at offset need-1 (exact dominance; a tie fails it)
        const uint32_t o = need >= 4 ? need - 4 : 0;
        const uint32_t mask = need >= 4 ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (8 * (4 - need)));
        const int a = xi + (int)o;                         // anchor: candidate d starts at a - d
        const uint32_t key = bytes_at(W, a, w_last) & mask;
        // second filter: a distinctive window inside the bytes already matched
        const int so = pick_second_window(W, xi, best, w_last, lane);
        const uint32_t key2 = so >= 0 ? bytes_at(W, xi + so, w_last) : 0u;
        const int delta = (int)o - so;                     // second window sits delta bytes before the first
        const int dq = delta >> 2, dr = delta & 3;
        const int back_words = dq + (dr ? 1 : 0), sh8 = dr ? 8 * (4 - dr) : 0;
        const int c_hi = a - (int)d0;                      // nearest candidate still open
        const int c_lo = a - (int)reach;                   // farthest candidate
        bool improved = false;
        // one step = 64 words = 256 candidates: every lane takes the word 32 below its first one as well,
        // so that the loop control, the vote and the range test are paid once per 256 candidates
        for (int wtop = c_hi >> 2; (wtop << 2) + 3 >= c_lo && !improved; wtop -= 64) {
            uint32_t hb0 = 0u, hb1 = 0u;                   // hits in the nearer word / in the farther word
            n_steps++;
            // interior steps: all 256 candidates of the warp lie strictly inside (c_lo, c_hi]
            const bool interior = (wtop << 2) + 3 <= c_hi && ((wtop - 63) << 2) >= c_lo;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int w = wtop - 32 * half - lane;     // lane 0 holds the nearest word of its half
                if (interior || (w << 2) + 3 >= c_lo) {
                    SQZ_CHECK(w >= 0 && w <= w_last, "finish: candidate word outside the image");
                    const uint32_t low = W[w], hiw = word_at(W, w + 1, w_last);
                    const uint32_t t0 = (low ^ key) & mask;
                    const uint32_t t1 = (__byte_perm(low, hiw, 0x4321) ^ key) & mask;
                    const uint32_t t2 = (__byte_perm(low, hiw, 0x5432) ^ key) & mask;
                    const uint32_t t3 = (__byte_perm(low, hiw, 0x6543) ^ key) & mask;
                    uint32_t hits = (t0 == 0 ? 1u : 0u) | (t1 == 0 ? 2u : 0u) | (t2 == 0 ? 4u : 0u) | (t3 == 0 ? 8u : 0u);
                    if (!interior) {
                        const int kmax = min(3, c_hi - (w << 2));  // only the very first word is cut at the top
                        const int kmin = max(0, c_lo - (w << 2));
                        hits &= kmax >= 0 ? (2u << kmax) - 1u : 0u;
                        hits &= ~((1u << kmin) - 1u);
                    }
                    if (hits != 0 && so >= 0) {
                        const int w2 = w - back_words;         // >= 0: the window lies inside the match
                        // words below the image start can only feed candidates that are cut off anyway
                        SQZ_CHECK(w2 <= w_last, "finish: second window outside the image");
                        const uint32_t x0 = w2 >= 0 ? W[w2] : 0u;
                        const uint32_t x1 = w2 + 1 >= 0 ? word_at(W, w2 + 1, w_last) : 0u;
                        const uint32_t x2 = w2 + 2 >= 0 ? word_at(W, w2 + 2, w_last) : 0u;
                        const uint32_t lo2 = fsr(x0, x1, sh8), hi2 = fsr(x1, x2, sh8);
                        const uint32_t u0 = lo2 ^ key2;
                        const uint32_t u1 = __byte_perm(lo2, hi2, 0x4321) ^ key2;
                        const uint32_t u2 = __byte_perm(lo2, hi2, 0x5432) ^ key2;
                        const uint32_t u3 = __byte_perm(lo2, hi2, 0x6543) ^ key2;
                        hits &= (u0 == 0 ? 1u : 0u) | (u1 == 0 ? 2u : 0u) | (u2 == 0 ? 4u : 0u) | (u3 == 0 ? 8u : 0u);
                    }
                    if (half == 0) { hb0 = hits; } else { hb1 = hits; }
                }
            }
            if (__ballot_sync(0xFFFFFFFFu, (hb0 | hb1) != 0) == 0) { continue; }
#pragma unroll
            for (int half = 0; half < 2; half++) {         // the nearer half first
                const int w = wtop - 32 * half - lane;
                uint32_t& mine = half == 0 ? hb0 : hb1;
                while (!improved) {
                    const uint32_t any = __ballot_sync(0xFFFFFFFFu, mine != 0);
                    if (any == 0) { break; }
                    const int src = __ffs((int)any) - 1;                       // lowest lane = nearest word
                    const int kk = 31 - __clz((int)(mine | 1u));               // nearest candidate in my word
                    const int c = __shfl_sync(0xFFFFFFFFu, (w << 2) + kk, src);
                    const uint32_t hit_d = (uint32_t)(a - c);
                    // cooperative verify: common prefix of S[xi..] and S[xi-hit_d..], capped at room
                    uint32_t m = room;
                    n_verify++;
                    for (uint32_t base = 0; base < room; base += 32) {
                        n_rounds++;
                        const uint32_t k = base + (uint32_t)lane;
                        SQZ_CHECK(k >= room || (xi - (int)hit_d + (int)k >= 0 && xi + (int)k < x_end), "finish: verify outside the image");
                        const bool diff = k < room && S[xi + (int)k] != S[xi - (int)hit_d + (int)k];
                        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, diff);
                        if (bal != 0) { m = base + (uint32_t)(__ffs((int)bal) - 1); break; }
                    }
                    if (m >= need) {
                        if (runner_up != nullptr) { *runner_up = best; }   // longest run among the nearer candidates
                        best = m; bdist = hit_d; improved = true; n_improve++;
                        break;
                    }
                    if (lane == src) { mine &= ~(1u << kk); }
                }
            }
        }
        if (!improved) { break; }
        d0 = bdist + 1;
    }
    if (dbg != nullptr && lane == 0) {
        atomicAdd(dbg + 0, 1ull); atomicAdd(dbg + 1, (unsigned long long)n_steps);
        atomicAdd(dbg + 2, (unsigned long long)n_verify); atomicAdd(dbg + 3, (unsigned long long)n_improve);
        atomicAdd(dbg + 4, (unsigned long long)n_rounds);
    }
}

// A position leaves phase 1: its table word keeps what was found so far, marked open, with the
// distance below which everything is settled.  No value comes back (RED, not a load + store); a
// fresh best may still have a nearer equal inside the current group, so phase 2 starts it over.
__device__ __forceinline__ void hand_over(uint32_t* slot, bool fresh, uint32_t resume_tag, uint32_t slice_tag) {
    if (fresh) { *slot = kOpenBit | slice_tag; }
    else       { atomicOr(slot, kOpenBit | resume_tag); }
}

// Small shards (fewer tiles than the device has room for) split the distance range as well:
// blockIdx.y = slice, a slice scans slice_words word distances (a multiple of 32, i.e. of 1024
// distances -- the unit of the resume tag) into a table of its own (slice 0 into the final table,
// slice k into slice_tables + (k-1) * slice_stride), and combine_slices folds them: nearest slice first, strictly longer wins, as
// if one scan had walked them in order.  A slice knows nothing of what nearer slices found, so it
// lets more candidates through, but a warp walks 1/S of the distances: what bounds a small input
// is the latency of one warp's walk, not throughput.  slice_words = 0: one scan over everything.
// To keep the far slices from treating every three-byte coincidence as an improvement (measured:
// 28 instead of 2.7 scalar-path candidates per position on ELF data), the nearest slice runs first
// and the others are seeded with its table (init_table): a position starts with the nearest
// slice's best as the run to beat -- strictly, the nearest slice is nearer -- and a position the
// nearest slice handed over is closed in every other slice, because phase 2 will walk on from
// there through all farther distances anyway.
template <int kMinLen, bool kEdge, int kQ>
__global__ void __launch_bounds__(kThreads, kQ == 4 ? 3 : 5)
match_table(const uint8_t* __restrict__ shard, long long back, long long n, long long ahead,
            uint32_t max_len, uint32_t max_dist, uint32_t* __restrict__ table,
            uint32_t* __restrict__ slice_tables, uint32_t* __restrict__ open_mask, int tile_first,
            int slice_words, long long slice_stride, int slice_base, const uint32_t* __restrict__ init_table,
            unsigned long long* __restrict__ tile_cycles) {
#ifdef SQZ_DEBUG_COUNTERS
    const long long t_begin = clock64();
#endif
    extern __shared__ __align__(16) uint8_t smem_raw[];
    constexpr int kWarpOwned = warp_owned(kQ), kTileBlocks = tile_blocks(kQ), kTilePos = tile_pos(kQ);
    constexpr int kStageBlocks = stage_blocks(kQ);
    constexpr int kGated = gated(kQ);
    constexpr int kQuietGroups = 4;                // quiet groups (of 128 distances) before the need masks are dropped
    constexpr int kQuietEntries = SQZ_QUIET_ENTRIES;   // a group is quiet when the warp's threads met at most this many candidates' iterations
    constexpr int kBusyEntries = SQZ_BUSY_ENTRIES;     // ... and brings the masks back when they met this many
    constexpr int kBusyIterations = SQZ_BUSY_ITERATIONS;             // masks off: iterations with a candidate in one group that bring them back at once
    const Geometry geo = geometry(max_len, max_dist, kEdge, init_table != nullptr, kQ);
    const uint4* PL = reinterpret_cast<const uint4*>(smem_raw);           // [plane_blocks][2]
    uint8_t* best_len = smem_raw + geo.region_bytes;                      // [kTilePos + 32]
    uint32_t* VL = reinterpret_cast<uint32_t*>(smem_raw + geo.region_bytes + kTilePos + 32);

    uint32_t* seed = VL + (kEdge ? geo.plane_blocks : 0);                  // [(kTileBlocks + 1)][kGated + 1], seeded slices only
    const int tile = tile_first + (int)blockIdx.x;
    const int slice = slice_base + (int)blockIdx.y;
    if (slice > 0) { table = slice_tables + (long long)(slice - 1) * slice_stride; }   // slice 0 writes the final table
    const long long tile_pos0 = (long long)tile * kTilePos;              // shard-relative
    const long long plane_pos0 = tile_pos0 - (long long)geo.back_blocks * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- stage: bytes -> bit planes (and validity), zero this tile's outputs ----
    // The window + tile arrive in pieces of kStageBlocks*32 bytes: 16-byte cp.async copies
    // (global -> shared, no registers) into a staging area -- the best_len bytes, unused until
    // the scan starts -- then warp ballots transpose every 32 bytes into 8 plane words.
    {
        uint32_t* PLw = reinterpret_cast<uint32_t*>(smem_raw);
        uint8_t* stage = best_len;
        const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
        const uintptr_t valid_lo = reinterpret_cast<uintptr_t>(shard) - (uintptr_t)back;
        const uintptr_t valid_hi = reinterpret_cast<uintptr_t>(shard) + (uintptr_t)(n + ahead);
        for (int b0 = 0; b0 < geo.plane_blocks; b0 += kStageBlocks) {
            const int nb = min(kStageBlocks, geo.plane_blocks - b0);
            const uint8_t* src0 = shard + plane_pos0 + 32LL * b0;
            const int mis = (int)(reinterpret_cast<uintptr_t>(src0) & 15);
            const uint8_t* base = src0 - mis;                   // 16-byte aligned
            const int chunks = (mis + nb * 32 + 15) >> 4;
            for (int c = threadIdx.x; c < chunks; c += kThreads) {
                const uint8_t* g = base + 16 * c;
                const uintptr_t ga = reinterpret_cast<uintptr_t>(g);
                if (ga >= valid_lo && ga + 16 <= valid_hi) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"
                                 :: "r"(stage_s + 16u * (uint32_t)c), "l"(g) : "memory");
                } else {                                        // chunk straddles the end of the data
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        stage[16 * c + k] = (ga + k >= valid_lo && ga + k < valid_hi) ? __ldg(g + k) : (uint8_t)0;
                    }
                }
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncthreads();
            for (int blk = b0 + warp; blk < b0 + nb; blk += kWarps) {
                const long long pos = plane_pos0 + (long long)blk * 32 + lane;
                const bool ok = pos >= -back && pos < n + ahead;
                const uint32_t byte = fold_byte(stage[mis + (blk - b0) * 32 + lane]);   // zero where there is no data
                uint32_t mine = 0;
#pragma unroll
                for (int bit = 0; bit < 8; bit++) {
                    const uint32_t w = __ballot_sync(0xFFFFFFFFu, (byte >> bit) & 1u);
                    if (lane == bit) { mine = w; }
                }
                if (lane < 8) { PLw[blk * 8 + lane] = mine; }
                if (kEdge) {
                    const uint32_t v = __ballot_sync(0xFFFFFFFFu, ok);
                    if (lane == 0) { VL[blk] = v; }
                }
            }
            __syncthreads();                                    // the next piece overwrites the staging area
        }
        for (int k = threadIdx.x; k < kTilePos + 32; k += kThreads) { best_len[k] = 0; }
        for (int k = threadIdx.x; k < kTilePos; k += kThreads) {
            const long long p = tile_pos0 + k;
            if (p < n) { table[p] = 0; }
        }
        if (init_table != nullptr) {
            // seeded slice: the nearest slice's result is the run to beat (state byte and need masks);
            // what it handed over is closed here
            for (int blk = warp; blk <= kTileBlocks; blk += kWarps) {
                const int k = blk * 32 + lane;
                const long long p = tile_pos0 + k;
                const uint32_t w = (blk < kTileBlocks && p < n) ? init_table[p] : 0u;
                const bool open = (w & kOpenBit) != 0;
                const uint32_t have = open ? 0u : min((w >> 16) & 0x3FFu, 31u);
                if (blk < kTileBlocks) { best_len[k] = open ? kHandOver : (uint8_t)have; }
#pragma unroll
                for (int g = 0; g < kGated; g++) {
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, have > (uint32_t)(kMinLen + g));
                    if (lane == g) { seed[blk * (kGated + 1) + g] = m; }
                }
                const uint32_t c = __ballot_sync(0xFFFFFFFFu, open);
                if (lane == kGated) { seed[blk * (kGated + 1) + kGated] = c; }
            }
        }
    }
    __syncthreads();

    // ---- phase 1: per-thread state ---------------------------------------------
    const int own0 = warp * kWarpOwned + lane * kQ;         // first block of this thread (tile-relative)
    const int blk0 = geo.back_blocks + own0;                // same, as plane block index
    uint32_t qv[kQ][8];
    uint32_t vq[kQ];
    // need = best+1 per position, bit-sliced as a thermometer: G[k] bit p set = byte offset
    // kMinLen+k has to match as well (need > kMinLen+k); closed = never a candidate again
    uint32_t G[kGated][kQ], closed_m[kQ];
#pragma unroll
    for (int q = 0; q < kQ; q++) {
        SQZ_CHECK(blk0 + q >= 0 && blk0 + q < geo.plane_blocks, "phase 1: query plane block out of range");
        load_planes(PL, blk0 + q, qv[q]);
        // closed from the start: the look-ahead block and positions past the shard
        const long long p0 = tile_pos0 + (long long)(own0 + q) * 32;
        uint32_t closed = 0;
        if (lane == 31 && q == kQ - 1) { closed = 0xFFFFFFFFu; }
        else if (p0 + 32 > n) { closed = p0 >= n ? 0xFFFFFFFFu : (0xFFFFFFFFu << (int)(n - p0)); }
        closed_m[q] = closed;
#pragma unroll
        for (int k = 0; k < kGated; k++) { G[k][q] = 0; }
        if (init_table != nullptr) {
#pragma unroll
            for (int k = 0; k < kGated; k++) { G[k][q] = seed[(own0 + q) * (kGated + 1) + k]; }
            closed_m[q] |= seed[(own0 + q) * (kGated + 1) + kGated];
        }
        vq[q] = kEdge ? VL[blk0 + q] : 0xFFFFFFFFu;
    }

    // farthest distance any position of this tile can use
    const long long tile_last = min(tile_pos0 + kTilePos - 1, n - 1);
    const uint32_t reach = (uint32_t)min((long long)max_dist, tile_last + back);
    int m_end = (int)((reach + 31) / 32);
    int m_begin = 1;
    if (slice_words > 0) {
        m_begin = 1 + slice * slice_words;
        m_end = min(m_end, m_begin + slice_words - 1);
    }
    const uint32_t tie_limit = slice_words > 0 ? kSlicedTieLimit : kTieLimit;
    // a fresh position handed over starts over in phase 2 -- from the first distance of this slice
    const uint32_t slice_tag = (uint32_t)min(31, (32 * (m_begin - 1) + 1) >> 10) << kResumeShift;

    // Distances are visited in groups of 128: for each bit offset sh the four distances
    // d_t = 32*(m0+t) - sh, t = 0..3, are evaluated together, because block q at d_t reads the
    // same shifted candidate word as block q+1 at d_{t+1}: 7 funnel shifts per plane serve 16
    // (block, distance) pairs.  Inside a group the order is not ascending, so a position whose
    // best is `fresh` (set inside the current group) keeps need = best, lets ties through, and
    // the scalar path prefers the nearer of two equal runs; at the end of the group fresh
    // positions are promoted to need = best+1.  Across groups distances ascend as in the
    // reference.
    uint32_t fresh[kQ];
#pragma unroll
    for (int q = 0; q < kQ; q++) { fresh[q] = 0; }
    bool gate_on = true;                    // need masks in use (decided per warp and group, see the end of the loop)
    // one word of warp-wide bookkeeping for that decision: bits 0-3 iterations of the current group, masks off,
    // in which some thread met a candidate; bits 4-10 consecutive quiet groups; bits 11-17 how many of those
    // it takes to drop the masks
    uint32_t quiet = (uint32_t)kQuietGroups << 11;
    uint32_t entered = 0;                   // iterations of the current group in which this thread met a candidate
    // debugging aid (tools/tile_cycles.py, -DSQZ_DEBUG_COUNTERS builds only): what the scalar path sees
#ifdef SQZ_DEBUG_COUNTERS
    unsigned int c_surv = 0, c_better = 0, c_tie_fresh = 0, c_reject = 0, c_hand = 0, c_iter_slow = 0, c_false = 0;
#define SQZ_COUNT(x) ((x)++)
#else
#define SQZ_COUNT(x) ((void)0)
#endif

    for (int m0 = m_begin; m0 <= m_end; m0 += kQ) {
        // raw candidate words j = 0..2*kQ-1 <-> plane block blk0 - m0 - (kQ-1) + j
        uint32_t cr[2 * kQ][8];
        uint32_t vr[2 * kQ];
        const int jb = blk0 - m0 - (kQ - 1);
        // a position handed over in this group (and not fresh) has every distance below the
        // group's first one settled: phase 2 resumes there (in units of 1024, rounded down)
        const uint32_t resume_tag = (uint32_t)min(31, (32 * (m0 - 1) + 1) >> 10) << kResumeShift;
#pragma unroll
        for (int j = 0; j < 2 * kQ; j++) {
            SQZ_CHECK(jb + j >= 0 && jb + j < geo.plane_blocks, "phase 1: candidate plane block out of range");
            load_planes(PL, jb + j, cr[j]);
            vr[j] = kEdge ? VL[jb + j] : 0xFFFFFFFFu;
        }
#pragma unroll 1
        for (int sh = 31; sh >= 0; sh--) {
            // e[q][t]: bit p set = byte p of block q differs from the byte d_t before it
            uint32_t e[kQ][kQ];
#pragma unroll
            for (int b = 0; b < 8; b++) {
                // quiet data: the hashed planes only (the same for the whole warp)
                if (b < kQuietPlanes || gate_on) {
                    uint32_t sw[2 * kQ - 1];
#pragma unroll
                    for (int j = 0; j < 2 * kQ - 1; j++) { sw[j] = fsr(cr[j][b], cr[j + 1][b], sh); }
#pragma unroll
                    for (int q = 0; q < kQ; q++) {
#pragma unroll
                        for (int t = 0; t < kQ; t++) {
                            if (b == 0) { e[q][t] = sw[q - t + kQ - 1] ^ qv[q][b]; }
                            else        { e[q][t] = or_xor(e[q][t], sw[q - t + kQ - 1], qv[q][b]); }
                        }
                    }
                }
            }
            if (kEdge) {
                uint32_t sv[2 * kQ - 1];
#pragma unroll
                for (int j = 0; j < 2 * kQ - 1; j++) { sv[j] = fsr(vr[j], vr[j + 1], sh); }
#pragma unroll
                for (int q = 0; q < kQ; q++) {
#pragma unroll
                    for (int t = 0; t < kQ; t++) { e[q][t] |= ~vq[q] | ~sv[q - t + kQ - 1]; }
                }
            }
            uint32_t en[kQ];                                  // look-ahead: first block of the next lane
#pragma unroll
            for (int t = 0; t < kQ; t++) { en[t] = __shfl_down_sync(0xFFFFFFFFu, e[0][t], 1); }
            uint32_t ib[kQ][kQ];
            uint32_t none = 0xFFFFFFFFu;
            if (gate_on) {                                    // (the same for the whole warp)
#pragma unroll
                for (int q = 0; q < kQ; q++) {
                    uint32_t all_t = 0xFFFFFFFFu;
#pragma unroll
                    for (int t = 0; t < kQ; t++) {
                        // a candidate fails at position p if any of the first need(p) bytes differs
                        const uint32_t lo = e[q][t], hi = q + 1 < kQ ? e[q + 1][t] : en[t];
                        uint32_t acc = lo | fsr(lo, hi, 1);
                        if (kMinLen >= 3) { acc |= fsr(lo, hi, 2); }
#pragma unroll
                        for (int k = 0; k < kGated; k++) { acc |= fsr(lo, hi, kMinLen + k) & G[k][q]; }
                        ib[q][t] = acc;
                        all_t &= acc;
                    }
                    none &= all_t | closed_m[q];              // closed positions never count
                }
            } else {
                // quiet data (an image): hardly any candidate matches even min_len bytes, so the need
                // masks filter nothing; test min_len bytes only, on the hashed planes only, and let the scalar
                // path judge the rest
#pragma unroll
                for (int q = 0; q < kQ; q++) {
                    uint32_t all_t = 0xFFFFFFFFu;
#pragma unroll
                    for (int t = 0; t < kQ; t++) {
                        const uint32_t lo = e[q][t], hi = q + 1 < kQ ? e[q + 1][t] : en[t];
                        uint32_t acc = lo | fsr(lo, hi, 1);
                        if (kMinLen >= 3) { acc |= fsr(lo, hi, 2); }
                        ib[q][t] = acc;
                        all_t &= acc;
                    }
                    none &= all_t | closed_m[q];
                }
                // a wrong guess must not last: the eighth iteration of a group in which some thread meets a
                // candidate brings the masks back at once (incompressible data has two or three such
                // iterations per group: candidates whose hashed bytes agree, see kQuietPlanes)
                if (__any_sync(0xFFFFFFFFu, none != 0xFFFFFFFFu) && (++quiet & 15u) == (uint32_t)kBusyIterations) {
                    gate_on = true;
                    quiet = (uint32_t)kQuietGroups << 11;     // no quiet groups
                }
            }
            if (none != 0xFFFFFFFFu) {
                // ---- scalar path: exact decision for the few surviving positions ----
                // The match words e[][] are not kept for this path (16 registers the loop has no room for):
                // a (block, distance) pair with a survivor rebuilds its two words from the plane words,
                // 16 instructions each.
                entered++;
                SQZ_COUNT(c_iter_slow);
                const int shx = sh;
#pragma unroll
                for (int q = 0; q < kQ; q++) {
                    if (kQ > 1) {                             // one test for the four distances of this block
                        if ((~(ib[q][0] & ib[q][1] & ib[q][kQ > 2 ? 2 : 0] & ib[q][kQ > 3 ? 3 : 0]) & ~closed_m[q]) == 0) { continue; }
                    }
#pragma unroll
                    for (int t = 0; t < kQ; t++) {            // ascending distance for every position
                        uint32_t todo = ~ib[q][t] & ~closed_m[q];       // closed since the masks were formed?
                        if (todo == 0) { continue; }
                        const uint32_t d = (uint32_t)(32 * (m0 + t) - sh);
                        if (d > reach) { continue; }
                        uint32_t lo = 0, hi = 0;
#pragma unroll
                        for (int b = 0; b < 8; b++) {
                            lo |= fsr(cr[q - t + kQ - 1][b], cr[q - t + kQ][b], shx) ^ qv[q][b];
                        }
                        if (kEdge) { lo |= ~vq[q] | ~fsr(vr[q - t + kQ - 1], vr[q - t + kQ], shx); }
                        if (q + 1 < kQ) {
#pragma unroll
                            for (int b = 0; b < 8; b++) {
                                hi |= fsr(cr[q + 1 - t + kQ - 1][b], cr[q + 1 - t + kQ][b], shx) ^ qv[q + 1][b];
                            }
                            if (kEdge) { hi |= ~vq[q + 1] | ~fsr(vr[q + 1 - t + kQ - 1], vr[q + 1 - t + kQ], shx); }
                        } else {
                            // the next lane's word: exact only if that lane compared all eight planes.  A warp in
                            // its quiet body compared the hashed ones; the others come from shared memory here.
                            hi = en[t];
                            if (kQuietPlanes < 8 && kQ > 1) { hi |= high_planes_differ(PL, blk0 + kQ, jb + 2 * kQ - 1 - t, shx); }
                        }
                        while (todo != 0) {
                            const int p = __ffs((int)todo) - 1;
                            todo &= todo - 1;
                            const uint32_t bit = 1u << p;
                            const int k = (own0 + q) * 32 + p;          // tile-relative position
                            SQZ_COUNT(c_surv);
                            SQZ_CHECK(k >= 0 && k < kTilePos && tile_pos0 + k < n, "phase 1: survivor outside the tile or the shard");
                            const uint32_t state = best_len[k];         // low 5 bits: best, high 3: near-ties seen
                            const uint32_t have = state & 31u;
                            const uint32_t win = fsr(lo, hi, p);        // differing bytes from position p on
                            uint32_t* slot = table + tile_pos0 + k;
                            if (win == 0) {
                                // at least 32 equal bytes: longer than the window, phase 2 finishes it.
                                // A fresh best may still have a nearer equal: let phase 2 start over.
                                best_len[k] = kHandOver;
                                hand_over(slot, (fresh[q] & bit) != 0, resume_tag, slice_tag);
                                closed_m[q] |= bit;
                                SQZ_COUNT(c_hand);
                                continue;
                            }
                            const uint32_t run = (uint32_t)(__ffs((int)win) - 1);
                            if (kQuietPlanes < 8 && run < (uint32_t)kMinLen) { SQZ_COUNT(c_false); continue; }   // equal hashes, different bytes
                            bool better = run > have;
                            if (run == have && (fresh[q] & bit)) { better = d < (*slot & 0xFFFFu); SQZ_COUNT(c_tie_fresh); }
                            if (better) {
                                SQZ_COUNT(c_better);
                                best_len[k] = (uint8_t)run;
                                *slot = (run << 16) | d;
                                // the first of an iteration's distances (t = 0) has nothing nearer left in the group
                                // -- a later candidate is nearer only if its t is smaller -- so its best is
                                // strict at once; the others stay fresh until the group ends
                                const bool strict_now = SQZ_T0_STRICT && t == 0;
                                if (strict_now) { fresh[q] &= ~bit; } else { fresh[q] |= bit; }
#pragma unroll
                                for (int g = 0; g < kGated; g++) {
                                    if (run + (strict_now ? 1u : 0u) > (uint32_t)(kMinLen + g)) { G[g][q] |= bit; }
                                }
                            } else if ((SQZ_COUNT(c_reject), (d & kTieMask) == 0) && gate_on) {
                                // a candidate that only ties or falls short although the need masks let it
                                // through: count a sample of them; a position that keeps attracting them is
                                // cheaper to finish in phase 2  (without the masks everything gets here: not counted)
                                if (state >= (tie_limit << 5)) {
                                    best_len[k] = kHandOver;
                                    hand_over(slot, (fresh[q] & bit) != 0, resume_tag, slice_tag);
                                    closed_m[q] |= bit;
                                } else if (state < (7u << 5)) {
                                    best_len[k] = (uint8_t)(state + 32u);
                                }
                            }
                        }
                    }
                }
            }
        }
        // end of the group: every smaller distance has been seen, fresh bests become strict
#pragma unroll
        for (int q = 0; q < kQ; q++) {
#pragma unroll
            for (int g = kGated - 1; g > 0; g--) { G[g][q] |= G[g - 1][q] & fresh[q]; }
            G[0][q] |= fresh[q];
            fresh[q] = 0;
        }
        // The need masks are a filter, not part of the decision: the warp drops them after four groups
        // (512 distances) in which its threads met next to no candidate, and takes them up again as soon
        // as a group brings more than a few.  The asymmetry is deliberate: without the masks, text makes
        // every thread meet candidates in every iteration, so a wrong guess must not last.
        {
            const uint32_t met = __reduce_add_sync(0xFFFFFFFFu, entered);   // thread-iterations with a candidate
#ifndef SQZ_GATE_ALWAYS
            if (kQ > 1) {       // (the latency shape keeps the masks and all eight planes: its loop is short, adapting measured slower)
                const uint32_t need = quiet >> 11;
                uint32_t groups = (quiet >> 4) & 127u;
                if (met >= (uint32_t)kBusyEntries) { groups = 0; } else if (met <= (uint32_t)kQuietEntries && groups < need) { groups++; }
                gate_on = groups < need;
                quiet = (need << 11) | (groups << 4);
            }
#else
            (void)met;
#endif
            entered = 0;
        }
    }
    // Work list of phase 2: one bit per position handed over, one word per block of 32 positions.
    // Every block of the shard is owned by exactly one thread, so the list needs no zeroing and no
    // atomics, and phase 2 reads n/8 bytes instead of scanning the table for marks.
#pragma unroll
    for (int q = 0; q < kQ; q++) {
        if (lane == 31 && q == kQ - 1) { continue; }               // look-ahead block: the next warp owns it
        const long long p0 = tile_pos0 + (long long)(own0 + q) * 32;
        if (p0 < n && open_mask != nullptr) {                      // (sliced launches: combine_slices writes the list)
            const uint32_t past_end = p0 + 32 > n ? (0xFFFFFFFFu << (int)(n - p0)) : 0u;
            open_mask[p0 >> 5] = closed_m[q] & ~past_end;
        }
    }
#ifdef SQZ_DEBUG_COUNTERS
    if (tile_cycles != nullptr) {          // debugging aid: per-tile duration and scalar-path counts
        unsigned long long* dbg2 = tile_cycles + (1 << 20) + 8;
        atomicAdd(dbg2 + 0, (unsigned long long)c_surv); atomicAdd(dbg2 + 1, (unsigned long long)c_better);
        atomicAdd(dbg2 + 2, (unsigned long long)c_tie_fresh); atomicAdd(dbg2 + 3, (unsigned long long)c_reject);
        atomicAdd(dbg2 + 4, (unsigned long long)c_hand); atomicAdd(dbg2 + 5, (unsigned long long)c_iter_slow);
        atomicAdd(dbg2 + 6, (unsigned long long)c_false);
        __syncthreads();
        if (threadIdx.x == 0) { tile_cycles[tile] = (unsigned long long)(clock64() - t_begin); }
    }
#else
    (void)tile_cycles;
#endif
}

// Fold the tables of a sliced launch into slice 0's (the final table) and write the work list.
// Slices are taken nearest first.  A position some slice left open stops the fold there: what
// that slice and the nearer ones have settled is the state phase 2 resumes from, at the open
// slice's resume distance (everything farther is searched again, exactly).
__global__ void __launch_bounds__(256)
combine_slices(uint32_t* __restrict__ table, const uint32_t* __restrict__ slice_tables, long long first, long long end,
               long long slice_stride, int slices, uint32_t* __restrict__ open_mask) {
    // positions [first, end): first is a multiple of 32, and the positions from `end` to the next multiple of 32
    // belong to nobody else (end is the end of the shard or a tile boundary)
    const long long padded = (end + 31) & ~31LL;
    for (long long p = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; p < padded; p += (long long)gridDim.x * blockDim.x) {
        uint32_t best = 0;
        bool open = false;
        if (p < end) {
            for (int sl = 0; sl < slices; sl++) {
                const uint32_t w = sl == 0 ? table[p] : slice_tables[(long long)(sl - 1) * slice_stride + p];
                if (w & kOpenBit) {
                    const uint32_t part = w & kStateMask;
                    if ((part >> 16) > (best >> 16)) { best = part; }
                    best |= kOpenBit | (w & (31u << kResumeShift));
                    open = true;
                    break;
                }
                if ((w >> 16) > (best >> 16)) { best = w; }
            }
            table[p] = best;
        }
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, open);
        if ((threadIdx.x & 31) == 0) { open_mask[p >> 5] = mask; }
    }
}

// ---------------------------------------------------------------------------
// phase 2 kernel.  Phase 1 left a work list: open_mask has one bit per position it handed over
// (the table word of such a position carries kOpenBit, what was found so far and the distance
// below which everything is settled).  A warp takes a segment of positions, reads the segment's
// mask words in one go and closes every listed position from the top down:
//
//   inheritance -- if position p+1 ended with (b', d'), b' < max_len, and byte p
//     equals byte p-d', then position p ends with exactly (b'+1, d'): no
//     candidate can give p more than b'+1 (it would give p+1 more than b'), and a
//     nearer one with b'+1 would have been p+1's nearest b'.  Inside a long match
//     every position but the last inherits, so a match costs one search, not one
//     per position.
//   capped inheritance -- if p+1 ended with (max_len, d') and byte p equals byte p-d', position
//     p reaches max_len at d' as well, and only a candidate nearer than d' with a full max_len
//     run can replace d' (the reference stops at the nearest one, squeeze.h:353).  Such a
//     candidate d'' has run(p+1, d'') = max_len-1 exactly.  So one search at p for the nearest
//     candidates below d' with a run of at least max_len-kSlack settles p, and tells how long
//     the longest nearer run m* is: the next max_len-1-m* positions below p cannot have a nearer
//     full run either (it would be a run > m* at p) and inherit (max_len, d') on a byte compare.
//     A block repeated at a large distance costs one search per ~kSlack positions instead of one
//     per position (BASELINE config 2, csrc.cat: 17 % of the positions).  d' = 1 needs no search.
//   search -- otherwise the exact warp-wide search (finish_position), bounded from above by the
//     neighbour's run + 1 when the neighbour is finished and not cut at max_len.
//
// counters[0] is the segment cursor.
// ---------------------------------------------------------------------------
constexpr int kChunk = 4096;              // positions per CTA work item of phase 2 (large shards)
constexpr uint32_t kSlack = 64;           // capped inheritance looks for nearer runs >= max_len - kSlack
constexpr int kFinishCtasPerSm = 6;       // resident CTAs the staged window allows (37 KB each at max_dist 32767)

// Phase 2 works on chunks: a CTA stages the chunk's bytes and their whole look-back window in
// shared memory once (max_dist + chunk + max_len bytes), then its four warps close the chunk's
// listed positions, one sub-segment (a quarter of the chunk or less) per warp at a time, from the
// top down.  A search step is a shared-memory load away instead of an L2 round trip, which is what
// bounds a small shard: a warp closes its positions one after the other, so the slowest warp sets
// the time.  Small shards get smaller chunks, enough of them for two waves of CTAs.
struct FinishShape { int chunk, sub, smem_bytes; };

inline FinishShape finish_shape(long long n, uint32_t max_len, uint32_t max_dist, int sms) {
    // eight chunks per resident CTA: where every position of a chunk needs a search (an ELF table),
    // the chunk's CTA is what the whole shard waits for
    long long chunk = n / (8LL * kFinishCtasPerSm * sms);
    chunk = chunk / 128 * 128;
    chunk = chunk < 128 ? 128 : (chunk > kChunk ? kChunk : chunk);
    FinishShape f;
    f.chunk = (int)chunk;
    // large shards: 512-position pieces (a boundary can cost a search inheritance would have saved);
    // small shards: 32-position pieces, the dense spots of an ELF table then spread over all warps
#ifndef SQZ_FINISH_SUB
#define SQZ_FINISH_SUB 512
#endif
    f.sub = f.chunk >= 2048 ? SQZ_FINISH_SUB : 32;
    f.chunk = f.chunk / f.sub * f.sub;                 // whole pieces
    f.smem_bytes = (int)(((long long)max_dist + chunk + max_len + 8 + 16 + 15) & ~15LL) + 16;
    return f;
}

__global__ void __launch_bounds__(kThreads, kFinishCtasPerSm)
finish_marked(const uint8_t* __restrict__ shard, long long back, long long n, long long ahead,
              uint32_t min_len, uint32_t max_len, uint32_t max_dist, uint32_t* __restrict__ table,
              const uint32_t* __restrict__ open_mask, unsigned int* __restrict__ counters,
              unsigned long long* __restrict__ dbg, int chunk, int sub) {
    extern __shared__ __align__(16) uint8_t img[];     // the chunk's bytes and window, 16-byte aligned like the source
    __shared__ unsigned int s_chunk, s_sub;
    const int lane = threadIdx.x & 31;
    const long long chunks = (n + chunk - 1) / chunk;
    const long long mask_words = (n + 31) >> 5;
    const int sub_words = sub >> 5;                    // sub is a multiple of 32, at most 1024
    // runs a nearer candidate must reach to matter for capped inheritance: never below 32, the
    // longest run phase 1 can have settled without handing the position over
    const uint32_t slack = max_len > 32 + kSlack ? kSlack : (max_len > 32 ? max_len - 32 : 0u);
    const uint32_t img_s = (uint32_t)__cvta_generic_to_shared(img);
    const uintptr_t valid_lo = reinterpret_cast<uintptr_t>(shard) - (uintptr_t)back;
    const uintptr_t valid_hi = reinterpret_cast<uintptr_t>(shard) + (uintptr_t)(n + ahead);
    for (;;) {
        __syncthreads();                               // everyone is done with the previous chunk's image
        if (threadIdx.x == 0) { s_chunk = atomicAdd(counters, 1u); s_sub = 0; }
        __syncthreads();
        const long long c = s_chunk;
        if (c >= chunks) { break; }
        const long long c0 = c * chunk;
        const long long c1 = min(c0 + chunk, n);
        // anything listed in this chunk?  (chunk / 32 <= 128 mask words, one per thread)
        const long long mw = (c0 >> 5) + threadIdx.x;
        const int listed_here = (threadIdx.x < (chunk >> 5) && mw < mask_words && open_mask[mw] != 0) ? 1 : 0;
        if (!__syncthreads_or(listed_here)) { continue; }
        // ---- stage [lo, hi) of the shard: 16-byte cp.async pieces, the ragged ends byte by byte ----
        const long long lo = c0 - min((long long)max_dist, c0 + back);
        const long long hi = min(c1 + (long long)max_len + 8, n + ahead);
        const uint8_t* src0 = shard + lo;
        const int mis = (int)(reinterpret_cast<uintptr_t>(src0) & 15);
        const uint8_t* base = src0 - mis;              // img[k] = base[k]; position x sits at img[x - lo + mis]
        const int pieces = (int)((mis + (hi - lo) + 15) >> 4) + 1;      // + one zeroed piece for word reads past the end
        for (int k = threadIdx.x; k < pieces; k += kThreads) {
            const uint8_t* g = base + 16 * k;
            const uintptr_t ga = reinterpret_cast<uintptr_t>(g);
            if (ga >= valid_lo && ga + 16 <= valid_hi) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(img_s + 16u * (uint32_t)k), "l"(g) : "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    img[16 * k + j] = (ga + j >= valid_lo && ga + j < valid_hi) ? __ldg(g + j) : (uint8_t)0;
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        const long long off = (long long)mis - lo;     // img[x + off] = byte at shard-relative position x (lo <= x < hi)
        // ---- the warps take sub-segments of the chunk until none is left ----
        for (;;) {
            unsigned int k = 0;
            if (lane == 0) { k = atomicAdd(&s_sub, 1u); }
            k = __shfl_sync(0xFFFFFFFFu, k, 0);
            const long long s0 = c0 + (long long)k * sub;
            if (s0 >= c1) { break; }
            const long long w0 = s0 >> 5;
            const long long w_end = min(mask_words, (c1 + 31) >> 5);        // the chunk's words only
            const uint32_t my_words = (lane < sub_words && w0 + lane < w_end) ? open_mask[w0 + lane] : 0u;
            uint32_t listed = __ballot_sync(0xFFFFFFFFu, my_words != 0);
            long long known_pos = -1;                      // position closed last by this warp ...
            uint32_t known_word = 0;                       // ... and its final word
            uint32_t cert = 0;                             // positions below known_pos that may inherit its capped result
            while (listed != 0) {
                const int wi = 31 - __clz((int)listed);    // highest block first
                listed &= ~(1u << wi);
                uint32_t marks = __shfl_sync(0xFFFFFFFFu, my_words, wi);
                const long long blk = s0 + 32LL * wi;
                while (marks != 0) {
                    const int src = 31 - __clz((int)marks);             // highest position first
                    marks &= ~(1u << src);
                    const long long p = blk + src;
                    SQZ_CHECK(p < c1, "phase 2: listed position outside the chunk");
                    const uint32_t raw = table[p];
                    SQZ_CHECK((raw & kOpenBit) != 0, "phase 2: listed position is not open");
                    const uint32_t word = raw & kStateMask;
                    const uint32_t resume = ((raw >> kResumeShift) & 31u) << 10;   // phase 1 settled every nearer distance
                    uint32_t best = word >> 16, bdist = word & 0xFFFFu;
                    const uint32_t room = (uint32_t)min((long long)max_len, n + ahead - p);
                    const uint32_t far = (uint32_t)min((long long)max_dist, p + back);
                    // the neighbour above: closed by this warp a moment ago, or read from the table
                    uint32_t nb = 0;
                    if (p + 1 < n) { nb = (p + 1 == known_pos) ? known_word : table[p + 1]; }
                    const uint32_t nlen = (nb >> 16) & 0x7FFFu, ndist = nb & 0xFFFFu;
                    const bool usable = (nb & kOpenBit) == 0 && nlen >= min_len && ndist >= 1 && ndist <= far;
                    const bool byte_same = usable && img[p + off] == img[p + off - (long long)ndist];
                    // 0 = closed without a search, 1 = full search, 2 = capped inheritance's search among nearer candidates
                    int mode = 1;
                    uint32_t cert_next = 0;
                    if (usable && nlen + 1 <= room && nlen < max_len) {
                        if (byte_same) { best = nlen + 1; bdist = ndist; mode = 0; }
                    } else if (byte_same && nlen == max_len && room == max_len) {
                        if (ndist == 1) {
                            // inside a run of one byte value: nothing is nearer than 1
                            best = max_len; bdist = 1; mode = 0;
                        } else if (cert > 0 && p + 1 == known_pos) {
                            // the search a few positions above saw no nearer run long enough to matter here
                            best = max_len; bdist = ndist; cert_next = cert - 1; mode = 0;
                        } else if (slack > 0) {
                            mode = 2;
                        }
                    }
                    if (mode == 1 && dbg != nullptr && lane == 0) {
                        // why a search: 16 neighbour open, 17 neighbour without a match, 18 neighbour at max_len (byte differs),
                        // 19 byte differs, 20 no neighbour in this shard, 21 other
                        int why = 21;
                        if (p + 1 >= n) { why = 20; }
                        else if (nb & kOpenBit) { why = 16; }
                        else if (nlen < min_len) { why = 17; }
                        else if (nlen >= max_len) { why = 18; }
                        else if (usable && nlen + 1 <= room) { why = 19; }
                        atomicAdd(dbg + why, 1ull);
                    }
                    if (mode != 0) {
                        // byte image of the search: starts at the farthest candidate, rounded down to a word
                        const long long first = p - (long long)far;
                        SQZ_CHECK(first >= lo && p + (long long)room <= hi, "phase 2: search window outside the staged image");
                        const int mis4 = (int)((mis + (first - lo)) & 3);
                        const long long left = n + ahead - p;
                        const int x_end = mis4 + (int)far + (int)min(left, (long long)max_len + 8);
                        uint32_t b2 = mode == 2 ? max_len - slack - 1 : best;
                        uint32_t d2 = mode == 2 ? 0u : bdist;
                        uint32_t runner_up = b2;
                        // A finished neighbour bounds this position from above: a candidate that gives p a run of L
                        // gives p+1 a run of L-1, so L <= (best run at p+1) + 1 -- unless that one was cut at max_len.
                        // The search ends at the nearest candidate that reaches the bound, and does not start at all
                        // when phase 1 had found one already.
                        uint32_t bound = room;
                        if (p + 1 < n && (nb & kOpenBit) == 0 && nlen != max_len) {
                            bound = min(room, (nlen >= min_len ? nlen : min_len - 1) + 1);
                        }
                        finish_position(img + (first + off - mis4), mis4 + (int)far, x_end, resume, mode == 2 ? ndist - 1 : far,
                                        bound, min_len, b2, d2, lane, dbg, &runner_up);
                        if (mode == 2) {
                            best = max_len;
                            bdist = b2 == max_len ? d2 : ndist;
                            cert_next = max_len - 1 - (b2 == max_len ? runner_up : b2);
                            if (dbg != nullptr && lane == 0) { atomicAdd(dbg + 22, 1ull); }
                        } else {
                            best = b2;
                            bdist = d2;
                        }
                    } else if (dbg != nullptr && lane == 0) {
                        atomicAdd(dbg + 5, 1ull);
                    }
                    cert = cert_next;
                    known_pos = p;
                    known_word = best >= min_len ? ((best << 16) | bdist) : 0u;
                    if (lane == 0) { table[p] = known_word; }
                }
            }
        }
    }
}

}  // namespace v2
