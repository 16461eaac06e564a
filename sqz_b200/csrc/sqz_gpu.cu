// sqz_gpu.cu -- B200 (sm_100a) LZ77 longest-match search + greedy parse of
// sqz-b200, and the C-ABI declared in include/sqz_gpu.h.
//
// What it replaces in the reference (leok7v/sqz, generation G1):
//   /root/reference/attic/map_experiment/squeeze.h:338-358  brute-force search
//   /root/reference/attic/map_experiment/squeeze.h:337,377-394  greedy parse
// Results are bit-exact: for every position the longest common prefix with a
// candidate at distance 1..max_dist (capped by max_len and the end of data),
// the nearest candidate among equals; then the orbit of position 0 under
// next[i] = i + (len[i] >= min_len ? len[i] : 1) as the token stream.
//
// There is no CPU fallback in this file: without a device every entry point
// returns ENODEV.
#include "sqz_gpu.h"

#include <cuda_runtime.h>

#include "match_bitsliced.cuh"
#include "lz_expand.cuh"

#include <atomic>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local char g_err[256] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char* what, cudaError_t ce = cudaSuccess) {
    if (ce != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(ce));
    } else {
        snprintf(g_err, sizeof(g_err), "%s", what);
    }
    return code;
}

static int cuda_code(cudaError_t ce) {
    switch (ce) {
        case cudaErrorMemoryAllocation: return ENOMEM;
        case cudaErrorNoDevice:
        case cudaErrorInsufficientDriver:
        case cudaErrorInvalidDevice: return ENODEV;
        default: return EIO;
    }
}

#define CU(call)                                                         \
    do {                                                                 \
        cudaError_t ce_ = (call);                                        \
        if (ce_ != cudaSuccess) { return fail(cuda_code(ce_), #call, ce_); } \
    } while (0)

#define LAUNCHED(name)                                                   \
    do {                                                                 \
        g_launches.fetch_add(1, std::memory_order_relaxed);              \
        cudaError_t ce_ = cudaGetLastError();                            \
        if (ce_ != cudaSuccess) { return fail(cuda_code(ce_), name, ce_); } \
    } while (0)

// ---------------------------------------------------------------------------
// kernel 1: match table, one position per thread, window staged in shared
// memory.  (First correct path; the tuned kernel replaces it behind the same
// launcher.)
//
// Shared-memory image: bytes [lo, hi) of the shard, where lo reaches back
// max_dist bytes before the tile and hi runs max_len bytes past it, shifted so
// that 16-byte global chunks land on 16-byte shared chunks.
//
// Scan order and acceptance are the reference's: distance 1 first, a candidate
// replaces the best only when strictly longer.  The filter that keeps the scan
// cheap is exact: to beat `best` a candidate must agree on the bytes
// [need-4, need) with need = max(best+1, min_len) (or on the first `need`
// bytes while need < 4); only candidates that pass are measured in full.
// ---------------------------------------------------------------------------
namespace v1 {

constexpr int kThreads = 512;

__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* S, int x) {
    const uint32_t* W = reinterpret_cast<const uint32_t*>(S);
    const int w = x >> 2;
    const uint32_t lo = W[w], hi = W[w + 1];
    return __funnelshift_r(lo, hi, (x & 3) * 8);
}

__global__ void __launch_bounds__(kThreads)
match_table(const uint8_t* __restrict__ shard, long long back, long long n,
            long long ahead, uint32_t min_len, uint32_t max_len,
            uint32_t max_dist, uint32_t* __restrict__ table) {
    extern __shared__ __align__(16) uint8_t S[];
    const long long tile0 = (long long)blockIdx.x * kThreads;
    const long long behind = min((long long)max_dist, tile0 + back);
    const long long lo = tile0 - behind;                        // first byte staged (shard-relative)
    const long long hi = min(tile0 + kThreads + (long long)max_len, n + ahead);
    const uint8_t* g = shard + lo;
    const int a = (int)(reinterpret_cast<uintptr_t>(g) & 15);   // S[a] holds byte `lo`
    const int span = a + (int)(hi - lo);                        // bytes of S in use
    const int chunks = (span + 8 + 15) >> 4;                    // + zero padding for word reads
    for (int c = threadIdx.x; c < chunks; c += kThreads) {
        const int s0 = c << 4;
        uint4 v;
        if (s0 >= a && s0 + 16 <= span) {
            v = __ldg(reinterpret_cast<const uint4*>(g - a + s0));
        } else {
            uint8_t b[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int s = s0 + k;
                b[k] = (s >= a && s < span) ? __ldg(g - a + s) : (uint8_t)0;
            }
            memcpy(&v, b, 16);
        }
        reinterpret_cast<uint4*>(S)[c] = v;
    }
    __syncthreads();

    const long long p = tile0 + threadIdx.x;
    if (p >= n) { return; }
    const int xi = a + (int)(p - lo);                           // S index of position p
    const uint32_t room = (uint32_t)min((long long)max_len, n + ahead - p);
    const uint32_t reach = (uint32_t)min((long long)max_dist, p + back);
    const uint32_t* W = reinterpret_cast<const uint32_t*>(S);

    uint32_t best = 0, bdist = 0;
    uint32_t d = 1;
    while (d <= reach) {
        const uint32_t need = max(best + 1, min_len);
        if (need > room) { break; }
        const uint32_t o = need >= 4 ? need - 4 : 0;
        const uint32_t mask = need >= 4 ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (8 * (4 - need)));
        const uint32_t key = load_u32_unaligned(S, xi + (int)o) & mask;
        int c = xi + (int)o - (int)d;                           // candidate word starts here
        const int c_end = xi + (int)o - (int)reach;             // farthest candidate
        int w = c >> 2;
        uint32_t hiw = W[w + 1], low = W[w];
        int kmax = c & 3;
        int found = -1;
        for (;;) {
            const uint32_t t3 = (__byte_perm(low, hiw, 0x6543) ^ key) & mask;
            const uint32_t t2 = (__byte_perm(low, hiw, 0x5432) ^ key) & mask;
            const uint32_t t1 = (__byte_perm(low, hiw, 0x4321) ^ key) & mask;
            const uint32_t t0 = (low ^ key) & mask;
            if (min(min(t0, t1), min(t2, t3)) == 0) {
                uint32_t hb = (t0 == 0 ? 1u : 0u) | (t1 == 0 ? 2u : 0u) |
                              (t2 == 0 ? 4u : 0u) | (t3 == 0 ? 8u : 0u);
                const int kmin = max(0, c_end - (w << 2));
                hb &= (2u << kmax) - 1u;
                hb &= ~((1u << kmin) - 1u);
                if (hb != 0) { found = (w << 2) + (31 - __clz(hb)); break; }
            }
            w--;
            if ((w << 2) + 3 < c_end) { break; }
            hiw = low;
            low = W[w];
            kmax = 3;
        }
        if (found < 0) { break; }
        const uint32_t dh = (uint32_t)(xi + (int)o - found);
        const uint8_t* A = S + xi;
        const uint8_t* B = S + xi - (int)dh;
        uint32_t m = 0;
        while (m < room && A[m] == B[m]) { m++; }
        if (m >= need) {
            best = m;
            bdist = dh;
            if (best >= room) { break; }
        }
        d = dh + 1;
    }
    table[p] = best >= min_len ? ((best << 16) | bdist) : 0u;
}

static size_t smem_bytes(uint32_t max_len, uint32_t max_dist) {
    size_t b = 16 + (size_t)max_dist + kThreads + max_len + 16 + 16;
    return (b + 15) & ~(size_t)15;
}

}  // namespace v1

// ---------------------------------------------------------------------------
// kernel 2: split packed table words into the two u16 arrays of the host ABI
// ---------------------------------------------------------------------------
__global__ void unpack_table(const uint32_t* __restrict__ table, size_t n,
                             uint16_t* __restrict__ len, uint16_t* __restrict__ dist) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint32_t w = table[i];
        len[i] = (uint16_t)(w >> 16);
        dist[i] = (uint16_t)(w & 0xFFFF);
    }
}

// ---------------------------------------------------------------------------
// greedy parse (squeeze.h:337,377-394) as block-wise pointer jumping.
//
// The shard is cut into parse blocks of kPB positions.  A token never jumps
// more than max_len, so a block can only be entered at an offset < max_len
// ("entry") and leaves with an overshoot < max_len into the next block.
//   parse_exit_map   per block: overshoot for every possible entry, by pointer
//                    doubling over next[] in shared memory
//   chain_groups     compose the maps of kGroup consecutive blocks
//   chain_top        serial walk over the groups -> entry of every group,
//                    overshoot of the whole shard
//   chain_expand     entry of every block
//   parse_walk<0>    per block (one warp, steps staged in shared memory): number of tokens on the real path
//   scan_counts      exclusive prefix sum -> output offsets, total
//   parse_walk<1>    per block: the same walk, then all lanes write the tokens (or symbol words)
// ---------------------------------------------------------------------------
namespace parse {

constexpr int kPBLarge = 4096;   // positions per parse block
constexpr int kPBSmall = 1024;   // ... of a small shard: the walks of a block are serial, shorter ones finish sooner
constexpr size_t kSmallShard = (size_t)8 << 20;
constexpr int kMapStride = 512;  // = sqz_gpu_max_len_limit entries per exit map
constexpr int kGroupMax = 128;   // blocks per chain group (large shards; small ones use about sqrt(blocks))

static int parse_block(size_t n) { return n <= kSmallShard ? kPBSmall : kPBLarge; }

static int group_size(size_t blocks) {
    int g = 8;
    while (g < kGroupMax && (size_t)g * g < blocks) { g <<= 1; }
    return g;
}

struct Work {                    // carved out of the caller's d_work
    uint16_t* exit_map;          // [blocks][kMapStride]
    uint16_t* group_map;         // [groups][kMapStride]
    uint32_t* group_entry;       // [groups + 1]
    uint32_t* block_entry;       // [blocks]
    uint32_t* count;             // [blocks]
    uint64_t* offset;            // [blocks]
    size_t blocks, groups;
    int pb, group;               // positions per block, blocks per group
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t workspace_for(size_t n, int pb) {
    const size_t blocks = (n + pb - 1) / pb + 1;
    const size_t groups = (blocks + 8 - 1) / 8 + 1;            // the smallest group size: an upper bound
    return align_up(blocks * kMapStride * 2) + align_up(groups * kMapStride * 2) +
           align_up((groups + 1) * 4) + align_up(blocks * 4) + align_up(blocks * 4) +
           align_up(blocks * 8) + 256;
}

// enough for every shard of at most n positions (small shards use smaller blocks)
static size_t workspace(size_t n) {
    return std::max(workspace_for(n, parse_block(n)), workspace_for(std::min(n, kSmallShard), kPBSmall));
}

static Work carve(void* d_work, size_t n) {
    Work w;
    w.pb = parse_block(n);
    w.blocks = (n + w.pb - 1) / w.pb;
    w.group = group_size(w.blocks);
    w.groups = (w.blocks + w.group - 1) / w.group;
    const size_t blocks = w.blocks + 1, groups = w.groups + 1;
    uint8_t* p = (uint8_t*)d_work;
    p = (uint8_t*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
    w.exit_map = (uint16_t*)p;    p += align_up(blocks * kMapStride * 2);
    w.group_map = (uint16_t*)p;   p += align_up(groups * kMapStride * 2);
    w.group_entry = (uint32_t*)p; p += align_up((groups + 1) * 4);
    w.block_entry = (uint32_t*)p; p += align_up(blocks * 4);
    w.count = (uint32_t*)p;       p += align_up(blocks * 4);
    w.offset = (uint64_t*)p;
    return w;
}

__device__ __forceinline__ uint32_t step_of(uint32_t word, uint32_t min_len) {
    const uint32_t len = word >> 16;
    return len >= min_len ? len : 1u;
}

template <int kPB>
__global__ void __launch_bounds__(256)
parse_exit_map(const uint32_t* __restrict__ table, size_t n, uint32_t min_len,
               uint32_t max_len, uint16_t* __restrict__ exit_map) {
    __shared__ uint16_t jump[kPB];
    const size_t b0 = (size_t)blockIdx.x * kPB;
    const int size = (int)min((size_t)kPB, n - b0);            // positions in this block
    for (int p = threadIdx.x; p < size; p += blockDim.x) {
        jump[p] = (uint16_t)(p + step_of(table[b0 + p], min_len));
    }
    __syncthreads();
    // asynchronous pointer doubling: every value stored is a point on p's own
    // path, so reading a neighbour's half-updated entry is still correct
    for (;;) {
        int changed = 0;
        for (int p = threadIdx.x; p < size; p += blockDim.x) {
            const int j = jump[p];
            if (j < size) {
                jump[p] = jump[j];
                changed = 1;
            }
        }
        if (!__syncthreads_or(changed)) { break; }
    }
    uint16_t* out = exit_map + (size_t)blockIdx.x * kMapStride;
    for (int e = threadIdx.x; e < (int)max_len; e += blockDim.x) {
        out[e] = (uint16_t)(e < size ? jump[e] - size : e - size);
    }
}

__global__ void chain_groups(const uint16_t* __restrict__ exit_map, size_t blocks, int group,
                             uint32_t max_len, uint16_t* __restrict__ group_map) {
    const size_t g = blockIdx.x;
    const size_t first = g * (size_t)group;
    const size_t last = min(first + (size_t)group, blocks);
    for (uint32_t e = threadIdx.x; e < max_len; e += blockDim.x) {
        uint32_t x = e;
        for (size_t b = first; b < last; b++) { x = exit_map[b * kMapStride + x]; }
        group_map[g * kMapStride + e] = (uint16_t)x;
    }
}

// one thread: entries of all groups, then the shard's overshoot and (later,
// from scan_counts) the token total
__global__ void chain_top(const uint16_t* __restrict__ group_map, size_t groups,
                          const uint32_t* __restrict__ entry_in, uint32_t entry_imm,
                          uint32_t* __restrict__ group_entry, uint64_t* __restrict__ result) {
    uint32_t x = entry_in != nullptr ? *entry_in : entry_imm;
    for (size_t g = 0; g < groups; g++) {
        group_entry[g] = x;
        x = group_map[g * kMapStride + x];
    }
    group_entry[groups] = x;
    result[1] = x;
}

// composed map of the whole shard, for seam hand-off between GPUs
__global__ void chain_total(const uint16_t* __restrict__ group_map, size_t groups,
                            uint32_t max_len, uint16_t* __restrict__ out) {
    for (uint32_t e = threadIdx.x; e < max_len; e += blockDim.x) {
        uint32_t x = e;
        for (size_t g = 0; g < groups; g++) { x = group_map[g * kMapStride + x]; }
        out[e] = (uint16_t)x;
    }
}

__global__ void chain_expand(const uint16_t* __restrict__ exit_map, size_t blocks, int group,
                             size_t groups, const uint32_t* __restrict__ group_entry,
                             uint32_t* __restrict__ block_entry) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) { return; }
    const size_t first = g * (size_t)group;
    const size_t last = min(first + (size_t)group, blocks);
    uint32_t x = group_entry[g];
    for (size_t b = first; b < last; b++) {
        block_entry[b] = x;
        x = exit_map[b * kMapStride + x];
    }
}

__global__ void __launch_bounds__(1024)
scan_counts(const uint32_t* __restrict__ count, size_t blocks,
            uint64_t* __restrict__ offset, uint64_t* __restrict__ result) {
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) { carry_s = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (size_t base = 0; base < blocks; base += 1024) {
        const size_t i = base + threadIdx.x;
        const uint64_t v = i < blocks ? count[i] : 0;
        uint64_t s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, s, d);
            if (lane >= d) { s += t; }
        }
        if (lane == 31) { warp_sum[wid] = s; }
        __syncthreads();
        if (wid == 0) {
            uint64_t ws = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, ws, d);
                if (lane >= d) { ws += t; }
            }
            warp_sum[lane] = ws;
        }
        __syncthreads();
        const uint64_t carry = carry_s;
        const uint64_t before = carry + (wid > 0 ? warp_sum[wid - 1] : 0) + s - v;
        if (i < blocks) { offset[i] = before; }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_s = before + v; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { result[0] = carry_s; }
}

// Symbol word of a match token (include/sqz_gpu.h): the bucket arithmetic of the reference's
// squeeze_encode_len / squeeze_encode_pos (squeeze.h:290-315, tables squeeze.h:29-79) done here,
// extra bits already in emission order (bit-reversed within their field).
//   lengths   3..10 -> buckets 0..7, no extra bits; above that four buckets per power of two of
//             len-3, with 258 kept in bucket 27 (squeeze.h:151-161)
//   distances 1..4 -> buckets 0..3; above that two buckets per power of two of dist-1
__device__ __forceinline__ uint32_t symbols_of_match(uint32_t len, uint32_t dist) {
    const uint32_t m = len - 3;
    uint32_t lb = m, lxb = 0;
    if (m >= 8) {
        const uint32_t k = 31u - (uint32_t)__clz((int)m);
        lxb = k - 2;
        lb = 4 * (k - 1) + ((m >> lxb) & 3);
    }
    const uint32_t lx = lxb ? __brev(m & ((1u << lxb) - 1)) >> (32 - lxb) : 0;
    const uint32_t q = dist - 1;
    uint32_t pb = q, pxb = 0;
    if (q >= 4) {
        const uint32_t k = 31u - (uint32_t)__clz((int)q);
        pxb = k - 1;
        pb = 2 * k + ((q >> pxb) & 1);
    }
    const uint32_t px = pxb ? __brev(q & ((1u << pxb) - 1)) >> (32 - pxb) : 0;
    return (257 + lb) | lx << 9 | pb << 14 | px << 19;
}

// The walk itself, one warp per block of kPB positions.  The warp stages the block's steps
// (len, or 1 for a literal) in shared memory with coalesced loads, lane 0 walks them from the
// block's entry offset -- about 30 cycles per token instead of a dependent global load -- and
// notes the positions it visits; then all lanes turn those positions into tokens (table word
// or literal byte, optionally as the coder's symbol word) and store them side by side.
// kEmit = false: count only (first pass, feeds the scan that places every block's tokens).
constexpr int kParseWarps = 2;

template <bool kEmit, bool kSymbols, int kPB>
__global__ void __launch_bounds__(32 * kParseWarps)
parse_walk(const uint8_t* __restrict__ shard, const uint32_t* __restrict__ table, size_t n,
           size_t blocks, uint32_t min_len, const uint32_t* __restrict__ block_entry,
           uint32_t* __restrict__ count, const uint64_t* __restrict__ offset,
           uint32_t* __restrict__ tokens, size_t cap) {
    __shared__ uint16_t step_s[kParseWarps][kPB];
    __shared__ uint16_t seen_s[kEmit ? kParseWarps : 1][kEmit ? kPB : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t b = (size_t)blockIdx.x * kParseWarps + warp;
    if (b >= blocks) { return; }                               // the whole warp leaves together
    const size_t b0 = b * kPB;
    const uint32_t size = (uint32_t)min((size_t)kPB, n - b0);
    for (uint32_t p = lane; p < size; p += 32) {
        step_s[warp][p] = (uint16_t)step_of(table[b0 + p], min_len);
    }
    __syncwarp();
    uint32_t c = 0;
    if (lane == 0) {
        uint32_t p = block_entry[b];
        while (p < size) {
            if (kEmit) { seen_s[warp][c] = (uint16_t)p; }
            p += step_s[warp][p];
            c++;
        }
    }
    c = __shfl_sync(0xFFFFFFFFu, c, 0);
    if (!kEmit) {
        if (lane == 0) { count[b] = c; }
        return;
    }
    __syncwarp();
    const uint64_t first = offset[b];
    for (uint32_t j = lane; j < c; j += 32) {
        const uint32_t p = seen_s[warp][j];
        const uint32_t w = table[b0 + p];
        const uint32_t len = w >> 16;
        uint32_t t;
        if (len >= min_len) { t = kSymbols ? symbols_of_match(len, w & 0xFFFF) : w; }
        else                { t = shard[b0 + p]; }
        if (first + j < cap) { tokens[first + j] = t; }
    }
}

}  // namespace parse

// ---------------------------------------------------------------------------
// launchers (device-buffer ABI)
// ---------------------------------------------------------------------------
static int check_rules(uint32_t min_len, uint32_t max_len, uint32_t max_dist) {
    if (min_len < 1 || max_len < min_len || max_len > sqz_gpu_max_len_limit ||
        max_dist < 1 || max_dist > sqz_gpu_max_dist_limit) {
        return fail(EINVAL, "rule parameters out of range (1 <= min_len <= max_len <= 512, 1 <= max_dist <= 65535)");
    }
    return 0;
}

// --- optional CUDA-event timing of the match kernel (bench.py roofline leg) ---
static std::mutex g_time_mu;
static bool g_timing = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_pending;
static double g_match_seconds = 0;
static uint64_t g_match_launches = 0;

static void timing_drain_locked() {
    for (auto& pr : g_pending) {
        float ms = 0;
        if (cudaEventSynchronize(pr.second) == cudaSuccess &&
            cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
            g_match_seconds += 1e-3 * (double)ms;
            g_match_launches++;
        }
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    g_pending.clear();
}

extern "C" void sqz_gpu_set_timing(int on) {
    std::lock_guard<std::mutex> lk(g_time_mu);
    g_timing = on != 0;
}

extern "C" double sqz_gpu_match_kernel_seconds(int reset, uint64_t* launches) {
    std::lock_guard<std::mutex> lk(g_time_mu);
    timing_drain_locked();
    const double avg = g_match_launches ? g_match_seconds / (double)g_match_launches : 0.0;
    if (launches != nullptr) { *launches = g_match_launches; }
    if (reset) { g_match_seconds = 0; g_match_launches = 0; }
    return avg;
}

static unsigned long long* g_tile_cycles = nullptr;   // debugging aid, see sqz_gpu_debug_tile_cycles

extern "C" void sqz_gpu_debug_tile_cycles(unsigned long long* d_buf) { g_tile_cycles = d_buf; }

// A/B switch for tests and measurements; per calling thread, so that it is no process-wide state
static thread_local int g_kernel_choice = 0;   // 0 auto, 1 thread-per-position (v1), 2 bit-sliced (v2)

extern "C" int sqz_gpu_select_kernel(int which) {
    if (which < 0 || which > 2) { return fail(EINVAL, "kernel choice must be 0, 1 or 2"); }
    g_kernel_choice = which;
    return 0;
}

// ---------------------------------------------------------------------------
// per-device set-up: function attributes, the two side streams of the edge tiles, SM count
// ---------------------------------------------------------------------------
constexpr int kMaxDevices = 64;

struct DeviceState {
    std::once_flag once;
    cudaError_t err = cudaSuccess;
    cudaStream_t side = nullptr, side2 = nullptr;
    int sms = 0;
};
static DeviceState g_dev[kMaxDevices];

template <typename K>
static cudaError_t allow_smem(K kernel, int bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    // several CTAs per SM only fit when the L1/shared split favours shared memory
    if (e == cudaSuccess) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    return e;
}

static int device_state(DeviceState** out) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) { return fail(ENODEV, "device index out of range"); }
    DeviceState& d = g_dev[dev];
    std::call_once(d.once, [&d, dev] {
        const int big4 = v2::geometry(sqz_gpu_max_len_limit, sqz_gpu_max_dist_limit, true, true, 4).smem_bytes;
        const int big1 = v2::geometry(sqz_gpu_max_len_limit, sqz_gpu_max_dist_limit, true, true, 1).smem_bytes;
        cudaError_t e = allow_smem(v2::match_table<3, false, 4>, big4);
        if (e == cudaSuccess) { e = allow_smem(v2::match_table<3, true, 4>, big4); }
        if (e == cudaSuccess) { e = allow_smem(v2::match_table<2, false, 4>, big4); }
        if (e == cudaSuccess) { e = allow_smem(v2::match_table<2, true, 4>, big4); }
        if (e == cudaSuccess) { e = allow_smem(v2::match_table<3, false, 1>, big1); }
        if (e == cudaSuccess) { e = allow_smem(v2::match_table<3, true, 1>, big1); }
        if (e == cudaSuccess) { e = allow_smem(v2::match_table<2, false, 1>, big1); }
        if (e == cudaSuccess) { e = allow_smem(v2::match_table<2, true, 1>, big1); }
        if (e == cudaSuccess) {
            e = cudaFuncSetAttribute(v1::match_table, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)v1::smem_bytes(sqz_gpu_max_len_limit, sqz_gpu_max_dist_limit));
        }
        if (e == cudaSuccess) {
            e = allow_smem(v2::finish_marked, v2::finish_shape(1LL << 30, sqz_gpu_max_len_limit, sqz_gpu_max_dist_limit, 148).smem_bytes);
        }
        if (e == cudaSuccess) { e = cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev); }
        if (e == cudaSuccess) { e = cudaStreamCreateWithFlags(&d.side, cudaStreamNonBlocking); }
        if (e == cudaSuccess) { e = cudaStreamCreateWithFlags(&d.side2, cudaStreamNonBlocking); }
        d.err = e;
    });
    if (d.err != cudaSuccess) { return fail(cuda_code(d.err), "per-device set-up", d.err); }
    *out = &d;
    return 0;
}

static int sm_count() {
    DeviceState* d = nullptr;
    return device_state(&d) == 0 ? d->sms : 148;
}

// an event that lives as long as one launcher call, whatever way the call ends
struct ScopedEvent {
    cudaEvent_t e = nullptr;
    ~ScopedEvent() { if (e != nullptr) { cudaEventDestroy(e); } }
    cudaError_t create() { return cudaEventCreateWithFlags(&e, cudaEventDisableTiming); }
};

// ---------------------------------------------------------------------------
// workspace of the match table: the segment cursor of phase 2 and phase 1's work list
// (one bit per position).  Callers of the _ws entry point own it; the plain entry point
// borrows one from a small per-device pool and hands it back when the stream has passed.
// ---------------------------------------------------------------------------
constexpr size_t kCursorBytes = 256;

// Distance slices of a small shard (match_bitsliced.cuh): as many as it takes to give the device
// about four waves of CTAs, a power of two, at most 32; none once the tiles alone fill it.
// Fixed numbers (444 resident CTAs: 148 SMs x 3), so that the workspace size depends on n only.
struct SlicePlan { int q; int slices; int slice_words; size_t stride; };

// How a shard is cut (match_bitsliced.cuh).  From a wave of throughput tiles on (444 resident CTAs of
// 16,256 positions: 148 SMs x 3) it is tiles only, four blocks per thread.  Below that the latency
// shape takes over -- one block per thread, 3,968 positions per CTA, five CTAs per SM -- and the
// distance range is split into as many slices as it takes to give the device about two waves of
// CTAs (a power of two, at most 32).  Fixed numbers, so that the workspace size depends on n only.
constexpr size_t kWaveQ4 = 148 * 3, kWaveQ1 = 148 * 5;

static SlicePlan slice_plan(size_t n, uint32_t max_dist) {
    SlicePlan sp{4, 1, 0, 0};
    const size_t tiles4 = (n + v2::tile_pos(4) - 1) / v2::tile_pos(4);
    int force_q = 0, force_s = 0;
#ifdef SQZ_TUNING
    if (const char* e = getenv("SQZ_Q")) { force_q = atoi(e); }
    if (const char* e = getenv("SQZ_SLICES")) { force_s = atoi(e); }
#endif
    if (n == 0 || (force_q == 0 && tiles4 > kWaveQ4) || force_q == 4) {
        if (force_s <= 1) { return sp; }
    } else {
        sp.q = 1;
    }
    const size_t tiles = (n + v2::tile_pos(sp.q) - 1) / v2::tile_pos(sp.q);
    const size_t want = force_s > 0 ? (size_t)force_s : (2 * (sp.q == 1 ? kWaveQ1 : kWaveQ4) + tiles - 1) / tiles;
    int slices = 1;
    while ((size_t)slices < want && slices < 32) { slices <<= 1; }
    const int words = (int)((max_dist + 31) / 32);                 // word distances of a full scan
    int per = (words + slices - 1) / slices;
    per = (per + 31) / 32 * 32;                                    // whole units of the resume tag (1024 distances)
    slices = (words + per - 1) / per;
    if (slices <= 1) { return sp; }
    sp.slices = slices;
    sp.slice_words = per;
    sp.stride = (n + 63) / 64 * 64;
    return sp;
}

static size_t mask_bytes(size_t n) { return (((n + 31) / 32 + 8) * 4 + 255) / 256 * 256; }

static size_t slice_bytes(const SlicePlan& sp) { return sp.slices > 1 ? (size_t)(sp.slices - 1) * sp.stride * 4 : 0; }

// A workspace sized for n serves every shard of at most n positions (the pipeline's slots are
// sized once and see chunks of many sizes): the slice tables are what a smaller shard may need
// more of, so their part is the largest any n' <= n asks for.
extern "C" size_t sqz_gpu_match_workspace(size_t n) {
    size_t tables = slice_bytes(slice_plan(n, sqz_gpu_max_dist_limit));
    const size_t tp = v2::tile_pos(1);
    const size_t tiles = std::min<size_t>((n + tp - 1) / tp, 2 * kWaveQ1 + 1);
    for (size_t t = 1; t <= tiles; t++) {
        tables = std::max(tables, slice_bytes(slice_plan(std::min(n, t * tp), sqz_gpu_max_dist_limit)));
    }
#ifdef SQZ_TUNING
    tables = std::max(tables, (size_t)31 * ((std::min<size_t>(n, (size_t)64 << 20) + 63) / 64 * 64) * 4);   // whatever SQZ_SLICES asks for
#endif
    return kCursorBytes + mask_bytes(n) + tables;
}

struct PoolBuffer { void* ptr; size_t bytes; cudaEvent_t passed; int device; };
static std::mutex g_pool_mu;
static std::vector<PoolBuffer> g_pool;          // idle or in flight (passed not yet reached)
constexpr size_t kPoolMax = 16;

static int pool_take(size_t bytes, PoolBuffer* out) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (size_t k = 0; k < g_pool.size(); k++) {
            PoolBuffer& b = g_pool[k];
            if (b.device == dev && b.bytes >= bytes && b.bytes <= 4 * bytes + (1u << 20) &&
                cudaEventQuery(b.passed) == cudaSuccess) {
                *out = b;
                g_pool.erase(g_pool.begin() + (long)k);
                return 0;
            }
        }
        cudaGetLastError();                      // cudaErrorNotReady of a busy buffer is not an error
    }
    PoolBuffer b{nullptr, bytes, nullptr, dev};
    CU(cudaMalloc(&b.ptr, bytes));
    cudaError_t ce = cudaEventCreateWithFlags(&b.passed, cudaEventDisableTiming);
    if (ce != cudaSuccess) { cudaFree(b.ptr); return fail(cuda_code(ce), "cudaEventCreate", ce); }
    *out = b;
    return 0;
}

static void pool_free(PoolBuffer& b) {
    cudaEventSynchronize(b.passed);
    cudaEventDestroy(b.passed);
    cudaFree(b.ptr);
}

// the work queued on `s` so far is the last user of the buffer
static void pool_give_back(PoolBuffer b, cudaStream_t s) {
    cudaEventRecord(b.passed, s);
    PoolBuffer victim{nullptr, 0, nullptr, 0};
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        g_pool.push_back(b);
        if (g_pool.size() > kPoolMax) { victim = g_pool.front(); g_pool.erase(g_pool.begin()); }
    }
    if (victim.ptr != nullptr) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(victim.device);
        pool_free(victim);
        cudaSetDevice(cur);
    }
}

static void pool_release_all() {
    std::vector<PoolBuffer> all;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        all.swap(g_pool);
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (PoolBuffer& b : all) { cudaSetDevice(b.device); pool_free(b); }
    cudaSetDevice(cur);
}

static int launch_v1(const uint8_t* d_shard, size_t back, size_t n, size_t ahead, uint32_t min_len,
                     uint32_t max_len, uint32_t max_dist, uint32_t* d_table, cudaStream_t s) {
    DeviceState* dv = nullptr;
    if (int r = device_state(&dv)) { return r; }
    const size_t tiles = (n + v1::kThreads - 1) / v1::kThreads;
    if (tiles > 0x7FFFFFFFull) { return fail(EINVAL, "shard too large for one launch"); }
    v1::match_table<<<(unsigned)tiles, v1::kThreads, v1::smem_bytes(max_len, max_dist), s>>>(
        d_shard, (long long)back, (long long)n, (long long)ahead, min_len, max_len, max_dist, d_table);
    LAUNCHED("match_table_v1");
    return 0;
}

template <int kMinLen, int kQ>
static int launch_tiles(const SlicePlan& sp, const uint8_t* d_shard, size_t back, size_t n, size_t ahead, uint32_t max_len,
                        uint32_t max_dist, uint32_t* d_table, void* d_work, cudaStream_t s) {
    DeviceState* dv = nullptr;
    if (int r = device_state(&dv)) { return r; }
    // Tiles whose every position sees the full max_dist window and max_len of
    // look-ahead run the plain variant; the rest (start of the first shard, end
    // of the last one, a partial last tile) run the variant with a validity plane.
    const long long tp = v2::tile_pos(kQ);
    const long long tiles = ((long long)n + tp - 1) / tp;
    if (tiles > 0x7FFFFFFFll) { return fail(EINVAL, "shard too large for one launch"); }
    long long t_lo = back >= max_dist ? 0 : ((long long)max_dist - (long long)back + tp - 1) / tp;
    long long tail = (long long)n + (long long)std::min<size_t>(ahead, max_len) - (long long)max_len;
    long long t_hi = tail <= 0 ? 0 : tail / tp;
    t_lo = std::min(t_lo, tiles);
    t_hi = std::max(std::min(t_hi, tiles), t_lo);
    if ((unsigned long long)n > 0xFFFFFFFFull) { return fail(EINVAL, "shard too large for one launch"); }
    unsigned int* d_counters = static_cast<unsigned int*>(d_work);
    uint32_t* d_open = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(d_work) + kCursorBytes);
    CU(cudaMemsetAsync(d_counters, 0, 16, s));
    const int smem_main = v2::geometry(max_len, max_dist, false, false, kQ).smem_bytes;
    const int smem_edge = v2::geometry(max_len, max_dist, true, false, kQ).smem_bytes;
    // small shards: the distance range is split across CTAs as well, the slices' tables sit behind
    // the work list and are folded into d_table afterwards
    uint32_t* d_slices = sp.slices > 1 ? reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(d_work) + kCursorBytes + mask_bytes(n)) : nullptr;
    // The few edge tiles run concurrently with the interior tiles: leading and trailing edge
    // tiles each on a side stream of the device, forked from and joined back into `s`.
    const bool lead = t_lo > 0, trail = tiles > t_hi;
    ScopedEvent fork, join1, join2;
    if (lead || trail) {
        CU(fork.create());
        CU(cudaEventRecord(fork.e, s));
    }
    // sliced: the nearest slice first (into d_table), then the others seeded with its result
    const int smem_main_seeded = v2::geometry(max_len, max_dist, false, true, kQ).smem_bytes;
    const int smem_edge_seeded = v2::geometry(max_len, max_dist, true, true, kQ).smem_bytes;
    auto launch = [&](auto kernel, long long first, long long count, bool edge, cudaStream_t on, int slices,
                      int slice_words, long long stride, uint32_t* slice_tables) -> cudaError_t {
        kernel<<<dim3((unsigned)count, 1), v2::kThreads, edge ? smem_edge : smem_main, on>>>(
            d_shard, (long long)back, (long long)n, (long long)ahead, max_len, max_dist, d_table, slice_tables,
            slices > 1 ? nullptr : d_open, (int)first, slice_words, stride, 0, nullptr, g_tile_cycles);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess || slices <= 1) { return e; }
        kernel<<<dim3((unsigned)count, (unsigned)(slices - 1)), v2::kThreads, edge ? smem_edge_seeded : smem_main_seeded, on>>>(
            d_shard, (long long)back, (long long)n, (long long)ahead, max_len, max_dist, d_table, slice_tables,
            nullptr, (int)first, slice_words, stride, 1, d_table, g_tile_cycles);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError();
    };
    cudaError_t le = cudaSuccess;
    if (lead) {
        CU(cudaStreamWaitEvent(dv->side, fork.e, 0));
        le = launch(v2::match_table<kMinLen, true, kQ>, 0, t_lo, true, dv->side, sp.slices, sp.slice_words, (long long)sp.stride, d_slices);
        if (le != cudaSuccess) { return fail(cuda_code(le), "match_table_v2_edge", le); }
        CU(join1.create());
        CU(cudaEventRecord(join1.e, dv->side));
    }
    if (trail) {
        CU(cudaStreamWaitEvent(dv->side2, fork.e, 0));
        le = launch(v2::match_table<kMinLen, true, kQ>, t_hi, tiles - t_hi, true, dv->side2, sp.slices, sp.slice_words, (long long)sp.stride, d_slices);
        if (le != cudaSuccess) { return fail(cuda_code(le), "match_table_v2_edge", le); }
        CU(join2.create());
        CU(cudaEventRecord(join2.e, dv->side2));
    }
    if (t_hi > t_lo) {
        le = launch(v2::match_table<kMinLen, false, kQ>, t_lo, t_hi - t_lo, false, s, sp.slices, sp.slice_words, (long long)sp.stride, d_slices);
        if (le != cudaSuccess) { return fail(cuda_code(le), "match_table_v2", le); }
    }
    if (join1.e != nullptr) { CU(cudaStreamWaitEvent(s, join1.e, 0)); }
    if (join2.e != nullptr) { CU(cudaStreamWaitEvent(s, join2.e, 0)); }
    if (sp.slices > 1) {
        const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)dv->sms * 8);
        v2::combine_slices<<<grid, 256, 0, s>>>(d_table, d_slices, 0LL, (long long)n, (long long)sp.stride, sp.slices, d_open);
        LAUNCHED("combine_slices");
    }
    const v2::FinishShape fs = v2::finish_shape((long long)n, max_len, max_dist, dv->sms);
    const long long chunks = ((long long)n + fs.chunk - 1) / fs.chunk;
    const unsigned finish_ctas = (unsigned)std::min<long long>(chunks, (long long)dv->sms * v2::kFinishCtasPerSm);
    v2::finish_marked<<<finish_ctas, v2::kThreads, fs.smem_bytes, s>>>(
        d_shard, (long long)back, (long long)n, (long long)ahead, (uint32_t)kMinLen, max_len, max_dist, d_table, d_open,
        d_counters, g_tile_cycles ? g_tile_cycles + (1 << 20) : nullptr, fs.chunk, fs.sub);
    LAUNCHED("match_finish_marked");
    return 0;
}

template <int kMinLen>
static int launch_v2(const uint8_t* d_shard, size_t back, size_t n, size_t ahead, uint32_t max_len,
                     uint32_t max_dist, uint32_t* d_table, void* d_work, cudaStream_t s) {
    const SlicePlan sp = slice_plan(n, max_dist);
    return sp.q == 1 ? launch_tiles<kMinLen, 1>(sp, d_shard, back, n, ahead, max_len, max_dist, d_table, d_work, s)
                     : launch_tiles<kMinLen, 4>(sp, d_shard, back, n, ahead, max_len, max_dist, d_table, d_work, s);
}

extern "C" int sqz_gpu_match_table_device_ws(const uint8_t* d_shard, size_t back, size_t n,
                                             size_t ahead, uint32_t min_len, uint32_t max_len,
                                             uint32_t max_dist, uint32_t* d_table, void* d_work, void* stream) {
    if (int r = check_rules(min_len, max_len, max_dist)) { return r; }
    if (n == 0) { return 0; }
    if (d_work == nullptr) { return fail(EINVAL, "null workspace"); }
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool timed = false;
    {
        std::lock_guard<std::mutex> lk(g_time_mu);
        timed = g_timing;
        if (timed && g_pending.size() > 4096) { timing_drain_locked(); }
    }
    if (timed) {
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, s));
    }
    const int choice = g_kernel_choice;
    int r;
    const bool v2_ok = max_len > 32;      // the bit-sliced kernel measures runs in a 32-bit window
    if (choice != 1 && v2_ok && min_len == 3)      { r = launch_v2<3>(d_shard, back, n, ahead, max_len, max_dist, d_table, d_work, s); }
    else if (choice != 1 && v2_ok && min_len == 2) { r = launch_v2<2>(d_shard, back, n, ahead, max_len, max_dist, d_table, d_work, s); }
    else { r = launch_v1(d_shard, back, n, ahead, min_len, max_len, max_dist, d_table, s); }
    if (timed) {
        if (r == 0 && cudaEventRecord(e1, s) == cudaSuccess) {
            std::lock_guard<std::mutex> lk(g_time_mu);
            g_pending.emplace_back(e0, e1);
        } else {
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
    }
    return r;
}

extern "C" int sqz_gpu_match_table_device(const uint8_t* d_shard, size_t back, size_t n,
                                          size_t ahead, uint32_t min_len, uint32_t max_len,
                                          uint32_t max_dist, uint32_t* d_table, void* stream) {
    if (int r = check_rules(min_len, max_len, max_dist)) { return r; }
    if (n == 0) { return 0; }
    PoolBuffer b;
    if (int r = pool_take(sqz_gpu_match_workspace(n), &b)) { return r; }
    const int r = sqz_gpu_match_table_device_ws(d_shard, back, n, ahead, min_len, max_len, max_dist, d_table, b.ptr, stream);
    pool_give_back(b, (cudaStream_t)stream);
    return r;
}

extern "C" int sqz_gpu_unpack_table_device(const uint32_t* d_table, size_t n,
                                           uint16_t* d_len, uint16_t* d_dist, void* stream) {
    if (n == 0) { return 0; }
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 16);
    unpack_table<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_table, n, d_len, d_dist);
    LAUNCHED("unpack_table");
    return 0;
}

extern "C" size_t sqz_gpu_parse_workspace(size_t n) { return parse::workspace(n); }

static int parse_maps(const uint32_t* d_table, size_t n, uint32_t min_len, uint32_t max_len,
                      const parse::Work& w, cudaStream_t s) {
    if (w.pb == parse::kPBSmall) {
        parse::parse_exit_map<parse::kPBSmall><<<(unsigned)w.blocks, 256, 0, s>>>(d_table, n, min_len, max_len, w.exit_map);
    } else {
        parse::parse_exit_map<parse::kPBLarge><<<(unsigned)w.blocks, 256, 0, s>>>(d_table, n, min_len, max_len, w.exit_map);
    }
    LAUNCHED("parse_exit_map");
    parse::chain_groups<<<(unsigned)w.groups, 256, 0, s>>>(w.exit_map, w.blocks, w.group, max_len, w.group_map);
    LAUNCHED("chain_groups");
    return 0;
}

// entry either from device memory (d_entry != null) or immediate
static int parse_launch(const uint8_t* d_shard, const uint32_t* d_table, size_t n,
                        const uint32_t* d_entry, uint32_t entry, uint32_t min_len,
                        uint32_t max_len, uint32_t* d_tokens, size_t cap, void* d_work,
                        uint64_t* d_result, cudaStream_t s, bool symbols = false) {
    if (n == 0) {
        // nothing to parse: count 0, overshoot = entry
        CU(cudaMemsetAsync(d_result, 0, 16, s));
        if (d_entry != nullptr) {
            CU(cudaMemcpyAsync(d_result + 1, d_entry, 4, cudaMemcpyDeviceToDevice, s));
        } else if (entry != 0) {
            return fail(EINVAL, "entry into an empty shard");
        }
        return 0;
    }
    parse::Work w = parse::carve(d_work, n);
    if (w.blocks > 0x7FFFFFFFull) { return fail(EINVAL, "shard too large for one launch"); }
    if (int r = parse_maps(d_table, n, min_len, max_len, w, s)) { return r; }
    parse::chain_top<<<1, 1, 0, s>>>(w.group_map, w.groups, d_entry, entry, w.group_entry, d_result);
    LAUNCHED("chain_top");
    parse::chain_expand<<<(unsigned)((w.groups + 127) / 128), 128, 0, s>>>(
        w.exit_map, w.blocks, w.group, w.groups, w.group_entry, w.block_entry);
    LAUNCHED("chain_expand");
    const unsigned wb = (unsigned)((w.blocks + parse::kParseWarps - 1) / parse::kParseWarps);
    const unsigned wt = 32 * parse::kParseWarps;
    const bool small = w.pb == parse::kPBSmall;
    if (small) {
        parse::parse_walk<false, false, parse::kPBSmall><<<wb, wt, 0, s>>>(d_shard, d_table, n, w.blocks, min_len, w.block_entry,
                                                                           w.count, nullptr, nullptr, 0);
    } else {
        parse::parse_walk<false, false, parse::kPBLarge><<<wb, wt, 0, s>>>(d_shard, d_table, n, w.blocks, min_len, w.block_entry,
                                                                           w.count, nullptr, nullptr, 0);
    }
    LAUNCHED("parse_count");
    parse::scan_counts<<<1, 1024, 0, s>>>(w.count, w.blocks, w.offset, d_result);
    LAUNCHED("scan_counts");
#define SQZ_EMIT(SYM, PB) parse::parse_walk<true, SYM, PB><<<wb, wt, 0, s>>>(d_shard, d_table, n, w.blocks, min_len, \
                                                                             w.block_entry, nullptr, w.offset, d_tokens, cap)
    if (symbols) { if (small) { SQZ_EMIT(true, parse::kPBSmall); } else { SQZ_EMIT(true, parse::kPBLarge); } }
    else         { if (small) { SQZ_EMIT(false, parse::kPBSmall); } else { SQZ_EMIT(false, parse::kPBLarge); } }
#undef SQZ_EMIT
    LAUNCHED("parse_emit");
    return 0;
}

extern "C" int sqz_gpu_parse_device(const uint8_t* d_shard, const uint32_t* d_table, size_t n,
                                    uint32_t entry, uint32_t min_len, uint32_t max_len,
                                    uint32_t* d_tokens, size_t tokens_cap, void* d_work,
                                    uint64_t* d_result, void* stream) {
    if (int r = check_rules(min_len, max_len, 1)) { return r; }
    if (entry >= max_len) { return fail(EINVAL, "entry must be < max_len"); }
    return parse_launch(d_shard, d_table, n, nullptr, entry, min_len, max_len, d_tokens,
                        tokens_cap, d_work, d_result, (cudaStream_t)stream);
}

// symbol words only exist for the bitstream's own limits (squeeze.h:13-15, 529-545)
static int check_symbol_rules(uint32_t min_len, uint32_t max_len, uint32_t max_dist) {
    // 258 has a bucket (27, extra bits 31) but no decoder accepts it: squeeze.h:529-545 and
    // sqz_decompress reject every length above 257 with EINVAL
    if (min_len < 3 || max_len > 257 || max_dist > 0x7FFF) {
        return fail(EINVAL, "symbol words need min_len >= 3, max_len <= 257, max_dist <= 32767");
    }
    return 0;
}

extern "C" int sqz_gpu_parse_symbols_device(const uint8_t* d_shard, const uint32_t* d_table, size_t n,
                                            uint32_t entry, uint32_t min_len, uint32_t max_len,
                                            uint32_t* d_words, size_t words_cap, void* d_work,
                                            uint64_t* d_result, void* stream) {
    if (int r = check_rules(min_len, max_len, 1)) { return r; }
    if (int r = check_symbol_rules(min_len, max_len, 1)) { return r; }
    if (entry >= max_len) { return fail(EINVAL, "entry must be < max_len"); }
    return parse_launch(d_shard, d_table, n, nullptr, entry, min_len, max_len, d_words,
                        words_cap, d_work, d_result, (cudaStream_t)stream, true);
}

extern "C" int sqz_gpu_parse_exit_map_device(const uint32_t* d_table, size_t n,
                                             uint32_t min_len, uint32_t max_len, void* d_work,
                                             uint16_t* d_exit_map, void* stream) {
    if (int r = check_rules(min_len, max_len, 1)) { return r; }
    if (n == 0) { return fail(EINVAL, "empty shard"); }
    cudaStream_t s = (cudaStream_t)stream;
    parse::Work w = parse::carve(d_work, n);
    if (int r = parse_maps(d_table, n, min_len, max_len, w, s)) { return r; }
    parse::chain_total<<<1, 512, 0, s>>>(w.group_map, w.groups, max_len, d_exit_map);
    LAUNCHED("chain_total");
    return 0;
}

// ---------------------------------------------------------------------------
// host-buffer ABI: chunked, double-buffered pipeline
// ---------------------------------------------------------------------------

// Tokens leave the device sized by their count, which only the device knows when the copy is
// queued: this kernel reads the count the parse left in result[0] and stores exactly that many
// words into pinned host memory (mapped into the device's address space by cudaHostAlloc) with
// 16-byte stores.  No host round trip, no copy of unused slots.
__global__ void __launch_bounds__(256)
tokens_to_host(const uint32_t* __restrict__ d_tokens, const uint64_t* __restrict__ d_result,
               uint32_t* __restrict__ h_tokens) {
    const size_t count = (size_t)d_result[0];
    const size_t quads = (count + 3) / 4;               // both buffers are 16-byte aligned and padded
    const uint4* src = reinterpret_cast<const uint4*>(d_tokens);
    uint4* dst = reinterpret_cast<uint4*>(h_tokens);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (size_t)gridDim.x * blockDim.x) {
        dst[i] = src[i];
    }
}

struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t ahead_of_all = nullptr;   // highest priority: a consumer's parse and token copy (kModeTokensPinned)
    cudaEvent_t done = nullptr;       // everything of the chunk has been issued and finished
    cudaEvent_t parsed = nullptr;     // the chunk's token count and overshoot are in h_result
    cudaEvent_t tabled = nullptr;     // the chunk's match table is complete
    uint8_t* d_data = nullptr;        // back halo + chunk + ahead halo
    uint32_t* d_table = nullptr;
    uint32_t* d_tokens = nullptr;
    uint16_t* d_len = nullptr;        // table mode only
    uint16_t* d_dist = nullptr;
    void* d_work = nullptr;           // parse workspace
    void* d_mwork = nullptr;          // match workspace (cursor + work list of phase 2)
    uint64_t* d_result = nullptr;     // [0] tokens, [1] overshoot (next chunk's entry)
    uint64_t* h_result = nullptr;     // pinned
    uint32_t* h_tokens = nullptr;     // pinned (streaming consumers only)
    size_t first = 0, n = 0;          // chunk = [first, first + n) of the input
    bool busy = false;
    // what the buffers were sized for (slots are recycled between calls, see slot_take)
    int device = -1;
    size_t cap_chunk = 0;
    uint32_t cap_len = 0, cap_dist = 0;
    int cap_mode = 0;
    size_t device_bytes = 0, pinned_bytes = 0;
};

enum { kModeTable = 0, kModeTokensPinned = 1, kModeTokensDirect = 2 };

struct sqz_gpu_stream {
    int device = 0;
    const uint8_t* data = nullptr;
    size_t bytes = 0;
    uint32_t min_len = 0, max_len = 0, max_dist = 0;
    size_t chunk = 0;
    size_t ramp = 0;                  // size of the next chunk while it is still below `chunk`
    size_t launched = 0;              // input bytes handed to the device so far
    size_t delivered = 0;             // input bytes whose tokens were returned
    int next_slot = 0, read_slot = 0;
    int mode = kModeTokensPinned;
    bool symbols = false;             // emit symbol words instead of plain tokens
    Slot slot[2];
    const uint64_t* prev_result = nullptr;   // device: previous chunk's result (entry hand-off)
    cudaEvent_t prev_parsed = nullptr;
    cudaEvent_t prev_tabled = nullptr;
};

// the calling thread's current device is the caller's business: every entry point that has to
// switch (an explicit stream device, the multi-device call) puts it back on the way out
struct DeviceScope {
    int saved = -1;
    DeviceScope() { if (cudaGetDevice(&saved) != cudaSuccess) { saved = -1; cudaGetLastError(); } }
    ~DeviceScope() { if (saved >= 0) { cudaSetDevice(saved); } }
};

static void slot_free(Slot& s) {
    if (s.stream) { cudaStreamSynchronize(s.stream); }
    cudaFree(s.d_data); cudaFree(s.d_table); cudaFree(s.d_tokens); cudaFree(s.d_len);
    cudaFree(s.d_dist); cudaFree(s.d_work); cudaFree(s.d_mwork); cudaFree(s.d_result);
    cudaFreeHost(s.h_result); cudaFreeHost(s.h_tokens);
    if (s.done) { cudaEventDestroy(s.done); }
    if (s.parsed) { cudaEventDestroy(s.parsed); }
    if (s.tabled) { cudaEventDestroy(s.tabled); }
    if (s.ahead_of_all) { cudaStreamSynchronize(s.ahead_of_all); cudaStreamDestroy(s.ahead_of_all); }
    if (s.stream) { cudaStreamDestroy(s.stream); }
    s = Slot();
}

static int slot_alloc(Slot& s, int device, size_t chunk, uint32_t max_len, uint32_t max_dist, int mode) {
    s.device = device; s.cap_chunk = chunk; s.cap_len = max_len; s.cap_dist = max_dist; s.cap_mode = mode;
    CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s.parsed, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s.tabled, cudaEventDisableTiming));
    const size_t data_bytes = (size_t)max_dist + chunk + max_len + 64;
    CU(cudaMalloc(&s.d_data, data_bytes));
    CU(cudaMalloc(&s.d_table, chunk * 4));
    CU(cudaMalloc(&s.d_mwork, sqz_gpu_match_workspace(chunk)));
    CU(cudaMalloc(&s.d_result, 16));
    CU(cudaHostAlloc(&s.h_result, 16, cudaHostAllocDefault));
    s.device_bytes = data_bytes + chunk * 4 + sqz_gpu_match_workspace(chunk);
    if (mode != kModeTable) {
        CU(cudaMalloc(&s.d_tokens, chunk * 4 + 16));
        CU(cudaMalloc(&s.d_work, parse::workspace(chunk)));
        s.device_bytes += chunk * 4 + parse::workspace(chunk);
        if (mode == kModeTokensPinned) {
            CU(cudaHostAlloc(&s.h_tokens, chunk * 4 + 16, cudaHostAllocDefault));
            s.pinned_bytes = chunk * 4 + 16;
            int least = 0, greatest = 0;
            CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            CU(cudaStreamCreateWithPriority(&s.ahead_of_all, cudaStreamNonBlocking, greatest));
        }
    } else {
        CU(cudaMalloc(&s.d_len, chunk * 2));
        CU(cudaMalloc(&s.d_dist, chunk * 2));
        s.device_bytes += chunk * 4;
    }
    return 0;
}

// Allocating and freeing gigabytes of device and pinned memory per call costs more than the
// search of a small input and stalls the device; finished calls park their slots here and the
// next call with the same shape takes them back.  At most kMaxParked slots and kMaxParkedBytes
// of device + pinned memory stay parked (oldest out first); sqz_gpu_release() frees them all.
static std::mutex g_park_mu;
static std::vector<Slot> g_parked;
constexpr size_t kMaxParked = 4;
constexpr size_t kMaxParkedBytes = (size_t)3 << 30;

static int slot_take(Slot& s, int device, size_t chunk, uint32_t max_len, uint32_t max_dist, int mode) {
    // round the capacity up to a power of two (>= 1 MiB) so that calls of similar size share slots
    size_t cap = (size_t)1 << 20;
    while (cap < chunk) { cap <<= 1; }
    chunk = cap;
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        for (size_t k = 0; k < g_parked.size(); k++) {
            const Slot& c = g_parked[k];
            if (c.device == device && c.cap_mode == mode && c.cap_chunk >= chunk &&
                c.cap_chunk <= 4 * chunk && c.cap_len >= max_len && c.cap_dist >= max_dist) {
                s = c;
                g_parked.erase(g_parked.begin() + (long)k);
                s.first = 0; s.n = 0; s.busy = false;
                return 0;
            }
        }
    }
    int r = slot_alloc(s, device, chunk, max_len, max_dist, mode);
    if (r != 0) { slot_free(s); }
    return r;
}

static void slot_give_back(Slot& s) {
    if (s.stream == nullptr) { s = Slot(); return; }
    cudaStreamSynchronize(s.stream);
    std::vector<Slot> victims;
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        g_parked.push_back(s);
        size_t total = 0;
        for (const Slot& c : g_parked) { total += c.device_bytes + c.pinned_bytes; }
        while (!g_parked.empty() && (g_parked.size() > kMaxParked || total > kMaxParkedBytes)) {
            total -= g_parked.front().device_bytes + g_parked.front().pinned_bytes;
            victims.push_back(g_parked.front());
            g_parked.erase(g_parked.begin());
        }
    }
    s = Slot();
    for (Slot& v : victims) { cudaSetDevice(v.device); slot_free(v); }
}

static void multi_release_all();

extern "C" void sqz_gpu_release(void) {
    std::vector<Slot> all;
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        all.swap(g_parked);
    }
    DeviceScope keep;
    for (Slot& c : all) { cudaSetDevice(c.device); slot_free(c); }
    pool_release_all();
    multi_release_all();
}

// device < 0 = the calling thread's current device
static int ensure_device(int& device) {
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0) {
        return fail(ENODEV, "no CUDA device: the match search has no CPU fallback", ce);
    }
    if (device < 0) { CU(cudaGetDevice(&device)); }
    if (device >= count) { return fail(ENODEV, "no such CUDA device"); }
    CU(cudaSetDevice(device));
    return 0;
}

// issue everything for the next chunk on its slot; returns without waiting
static int stream_launch_next(sqz_gpu_stream* st, uint16_t* len_out, uint16_t* dist_out) {
    Slot& s = st->slot[st->next_slot];
    const size_t first = st->launched;
    size_t n = std::min(st->chunk, st->bytes - first);
    if (st->ramp != 0) {               // short chunks first: the consumer starts after milliseconds
        n = std::min(n, st->ramp);
        st->ramp = st->ramp * 2 >= st->chunk ? 0 : st->ramp * 2;
    }
    const size_t back = std::min<size_t>(first, st->max_dist);
    const size_t ahead = std::min<size_t>(st->bytes - (first + n), st->max_len);
    s.first = first;
    s.n = n;
    CU(cudaMemcpyAsync(s.d_data, st->data + first - back, back + n + ahead,
                       cudaMemcpyHostToDevice, s.stream));
    const uint8_t* d_shard = s.d_data + back;
    if (st->mode == kModeTokensPinned && st->prev_tabled != nullptr) {
        // A consumer waits for the chunks in order: two searches sharing the device would both be ready
        // late.  This chunk's search starts when the one before has its table; only that one's parse and
        // copy run beside it.
        CU(cudaStreamWaitEvent(s.stream, st->prev_tabled, 0));
    }
    if (int r = sqz_gpu_match_table_device_ws(d_shard, back, n, ahead, st->min_len, st->max_len,
                                              st->max_dist, s.d_table, s.d_mwork, s.stream)) { return r; }
    CU(cudaEventRecord(s.tabled, s.stream));
    st->prev_tabled = s.tabled;
    // A consumer's parse and token copy go on a stream of the highest priority: the next chunk's search
    // is queued behind this one's and fills the device as soon as it may -- at equal priority the few
    // small kernels that finish this chunk would wait for free CTA slots until that search is over.
    cudaStream_t late = st->mode == kModeTokensPinned ? s.ahead_of_all : s.stream;
    if (late != s.stream) { CU(cudaStreamWaitEvent(late, s.tabled, 0)); }
    if (st->mode != kModeTable) {
        const uint32_t* d_entry = nullptr;
        if (st->prev_result != nullptr) {
            CU(cudaStreamWaitEvent(late, st->prev_parsed, 0));
            d_entry = reinterpret_cast<const uint32_t*>(st->prev_result + 1);  // low half of overshoot
        }
        if (int r = parse_launch(d_shard, s.d_table, n, d_entry, 0, st->min_len, st->max_len,
                                 s.d_tokens, n, s.d_work, s.d_result, late, st->symbols)) { return r; }
        CU(cudaMemcpyAsync(s.h_result, s.d_result, 16, cudaMemcpyDeviceToHost, late));
        CU(cudaEventRecord(s.parsed, late));
        st->prev_parsed = s.parsed;
        st->prev_result = s.d_result;
        if (st->mode == kModeTokensPinned) {
            // the tokens follow right away, exactly as many as there are: the consumer finds them
            // in pinned memory when it asks
            const unsigned grid = (unsigned)std::min<size_t>((n / 4 + 255) / 256 + 1, (size_t)sm_count() * 4);
            tokens_to_host<<<grid, 256, 0, late>>>(s.d_tokens, s.d_result, s.h_tokens);
            LAUNCHED("tokens_to_host");
        }
        // kModeTokensDirect: the caller copies count x 4 bytes to their final place once it knows the count
    } else {
        if (int r = sqz_gpu_unpack_table_device(s.d_table, n, s.d_len, s.d_dist, s.stream)) { return r; }
        CU(cudaMemcpyAsync(len_out + first, s.d_len, n * 2, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaMemcpyAsync(dist_out + first, s.d_dist, n * 2, cudaMemcpyDeviceToHost, s.stream));
    }
    if (late != s.stream) {             // the slot is free when both streams are through
        CU(cudaEventRecord(s.done, late));
        CU(cudaStreamWaitEvent(s.stream, s.done, 0));
    }
    CU(cudaEventRecord(s.done, s.stream));
    s.busy = true;
    st->launched = first + n;
    st->next_slot ^= 1;
    return 0;
}

// Token mode feeds a host consumer chunk by chunk: smaller chunks overlap better.
// Table mode only streams: larger chunks waste less on the last wave of tiles.
static size_t default_chunk(size_t bytes, bool consumer) {
    const size_t kDefault = consumer ? (size_t)32 << 20 : (size_t)128 << 20;
    return std::max<size_t>(std::min(bytes, kDefault), 1);
}

static int stream_open(sqz_gpu_stream** out, int device, const uint8_t* data, size_t bytes,
                       uint32_t window, uint32_t min_len, uint32_t max_len, uint32_t max_dist,
                       size_t chunk, int mode, bool short_start = false) {
    *out = nullptr;
    if (int r = check_rules(min_len, max_len, max_dist)) { return r; }
    if (window == 0 || (window & (window - 1)) != 0 || max_dist > window) {
        return fail(EINVAL, "window must be a power of two and max_dist <= window");
    }
    if (data == nullptr && bytes != 0) { return fail(EINVAL, "null input"); }
    if (int r = ensure_device(device)) { return r; }
    sqz_gpu_stream* st = new (std::nothrow) sqz_gpu_stream();
    if (st == nullptr) { return fail(ENOMEM, "out of host memory"); }
    st->device = device;
    st->data = data;
    st->bytes = bytes;
    st->min_len = min_len; st->max_len = max_len; st->max_dist = max_dist;
    st->chunk = chunk ? chunk : default_chunk(bytes, mode == kModeTokensPinned);
    st->mode = mode;
    // token streams with the default chunking start with short chunks: 2, 4, 8, 16 MiB, then 32 MiB
    if (mode == kModeTokensPinned && chunk == 0 && st->chunk > ((size_t)2 << 20)) { st->ramp = (size_t)2 << 20; }
    if (short_start && chunk >= ((size_t)4 << 20)) { st->ramp = chunk / 4; }
    const int slots = bytes > (st->ramp ? st->ramp : st->chunk) ? 2 : 1;
    for (int k = 0; k < slots && bytes > 0; k++) {
        if (int r = slot_take(st->slot[k], device, st->chunk, max_len, max_dist, mode)) {
            sqz_gpu_stream_close(st);
            return r;
        }
    }
    *out = st;
    return 0;
}

extern "C" int sqz_gpu_stream_open(sqz_gpu_stream** st, int device, const uint8_t* data,
                                   size_t bytes, uint32_t window, uint32_t min_len,
                                   uint32_t max_len, uint32_t max_dist, size_t chunk_bytes,
                                   uint32_t flags) {
    if (st == nullptr) { return fail(EINVAL, "null stream handle"); }
    if ((flags & ~(uint32_t)(SQZ_GPU_STREAM_SYMBOLS | SQZ_GPU_STREAM_SHORT_START)) != 0) {
        return fail(EINVAL, "unknown stream flags");
    }
    if (flags & SQZ_GPU_STREAM_SYMBOLS) {
        if (int r = check_symbol_rules(min_len, max_len, max_dist)) { return r; }
    }
    DeviceScope keep;
    if (int r = stream_open(st, device, data, bytes, window, min_len, max_len, max_dist,
                            chunk_bytes, kModeTokensPinned, (flags & SQZ_GPU_STREAM_SHORT_START) != 0)) { return r; }
    (*st)->symbols = (flags & SQZ_GPU_STREAM_SYMBOLS) != 0;
    if (bytes > 0) {
        if (int r = stream_launch_next(*st, nullptr, nullptr)) {
            sqz_gpu_stream_close(*st);
            *st = nullptr;
            return r;
        }
    }
    return 0;
}

extern "C" int sqz_gpu_stream_next(sqz_gpu_stream* st, const uint32_t** tokens, size_t* count) {
    if (st == nullptr || tokens == nullptr || count == nullptr) { return fail(EINVAL, "null argument"); }
    *tokens = nullptr;
    *count = 0;
    if (st->delivered >= st->bytes) { return 0; }
    DeviceScope keep;
    CU(cudaSetDevice(st->device));
    // keep the device busy: the slot the caller has just finished reading is free now
    static const bool trace = getenv("SQZ_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    if (st->launched < st->bytes) {
        if (int r = stream_launch_next(st, nullptr, nullptr)) { return r; }
    }
    const auto t1 = std::chrono::steady_clock::now();
    Slot& s = st->slot[st->read_slot];
    CU(cudaEventSynchronize(s.done));
    if (trace) {
        const auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "sqz_gpu_stream_next: queued the next chunk in %.1f ms, waited %.1f ms for this one\n",
                std::chrono::duration<double, std::milli>(t1 - t0).count(),
                std::chrono::duration<double, std::milli>(t2 - t1).count());
    }
    const uint64_t n_tok = s.h_result[0];
    if (n_tok > s.n) { return fail(EIO, "parse produced more tokens than positions"); }
    s.busy = false;
    st->delivered = s.first + s.n;
    st->read_slot ^= 1;
    *tokens = s.h_tokens;
    *count = (size_t)n_tok;
    return 0;
}

extern "C" void sqz_gpu_stream_close(sqz_gpu_stream* st) {
    if (st == nullptr) { return; }
    DeviceScope keep;
    cudaSetDevice(st->device);
    slot_give_back(st->slot[0]);
    slot_give_back(st->slot[1]);
    delete st;
}

// One-shot token call: the chunks' tokens go straight from device memory to their final place in
// the caller's buffer, count x 4 bytes each (a pinned destination takes them by DMA while the next
// chunk is searched; a pageable one through the driver's staging).  Nothing consumes the chunks
// on the way, so the larger streaming chunk is used.
extern "C" int sqz_gpu_tokens(const uint8_t* data, size_t bytes, uint32_t window,
                              uint32_t min_len, uint32_t max_len, uint32_t max_dist,
                              uint32_t* tokens_out, size_t tokens_cap, size_t* n_tokens) {
    if (n_tokens == nullptr) { return fail(EINVAL, "null n_tokens"); }
    *n_tokens = 0;
    sqz_gpu_stream* st = nullptr;
    DeviceScope keep;
    if (int r = stream_open(&st, -1, data, bytes, window, min_len, max_len, max_dist, 0, kModeTokensDirect)) {
        return r;
    }
    size_t total = 0;
    int rc = 0;
    if (bytes > 0) { rc = stream_launch_next(st, nullptr, nullptr); }
    while (rc == 0 && st->delivered < st->bytes) {
        Slot& s = st->slot[st->read_slot];
        // the other slot is free (its copy was queued before its `done`): keep the device busy
        if (st->launched < st->bytes) {
            Slot& o = st->slot[st->next_slot];
            if (o.busy) {
                cudaError_t ce = cudaEventSynchronize(o.done);
                if (ce != cudaSuccess) { rc = fail(cuda_code(ce), "cudaEventSynchronize", ce); break; }
                o.busy = false;
            }
            rc = stream_launch_next(st, nullptr, nullptr);
            if (rc != 0) { break; }
        }
        cudaError_t ce = cudaEventSynchronize(s.parsed);
        if (ce != cudaSuccess) { rc = fail(cuda_code(ce), "cudaEventSynchronize", ce); break; }
        const size_t c = (size_t)s.h_result[0];
        if (c > s.n) { rc = fail(EIO, "parse produced more tokens than positions"); break; }
        if (tokens_out != nullptr && total < tokens_cap) {
            ce = cudaMemcpyAsync(tokens_out + total, s.d_tokens, std::min(c, tokens_cap - total) * 4,
                                 cudaMemcpyDeviceToHost, s.stream);
            if (ce != cudaSuccess) { rc = fail(cuda_code(ce), "cudaMemcpyAsync", ce); break; }
        }
        ce = cudaEventRecord(s.done, s.stream);      // the slot is free again once the copy has landed
        if (ce != cudaSuccess) { rc = fail(cuda_code(ce), "cudaEventRecord", ce); break; }
        total += c;
        st->delivered = s.first + s.n;
        st->read_slot ^= 1;
    }
    for (int k = 0; k < 2; k++) {
        if (st->slot[k].stream != nullptr) {
            cudaError_t ce = cudaStreamSynchronize(st->slot[k].stream);
            if (ce != cudaSuccess && rc == 0) { rc = fail(cuda_code(ce), "cudaStreamSynchronize", ce); }
        }
    }
    sqz_gpu_stream_close(st);
    *n_tokens = total;
    if (rc == 0 && total > tokens_cap) { rc = fail(E2BIG, "token buffer too small"); }
    return rc;
}

extern "C" int sqz_gpu_match_table(const uint8_t* data, size_t bytes, uint32_t window,
                                   uint32_t min_len, uint32_t max_len, uint32_t max_dist,
                                   uint16_t* len_out, uint16_t* dist_out) {
    if (bytes != 0 && (len_out == nullptr || dist_out == nullptr)) {
        return fail(EINVAL, "null output");
    }
    sqz_gpu_stream* st = nullptr;
    DeviceScope keep;
    if (int r = stream_open(&st, -1, data, bytes, window, min_len, max_len, max_dist, 0, kModeTable)) {
        return r;
    }
    int rc = 0;
    while (rc == 0 && st->launched < st->bytes) {
        Slot& s = st->slot[st->next_slot];
        if (s.busy) {                      // reuse only after its copies have landed
            cudaError_t ce = cudaEventSynchronize(s.done);
            if (ce != cudaSuccess) { rc = fail(cuda_code(ce), "cudaEventSynchronize", ce); break; }
            s.busy = false;
        }
        rc = stream_launch_next(st, len_out, dist_out);
    }
    for (int k = 0; k < 2; k++) {
        if (st->slot[k].stream != nullptr) {
            cudaError_t ce = cudaStreamSynchronize(st->slot[k].stream);
            if (ce != cudaSuccess && rc == 0) { rc = fail(cuda_code(ce), "cudaStreamSynchronize", ce); }
        }
    }
    sqz_gpu_stream_close(st);
    return rc;
}

// ---------------------------------------------------------------------------
// several devices, one call.  The input is cut into contiguous shards,
// one per device, each uploaded with its look-back and look-ahead halo, so the match tables need
// no exchange at all.  The greedy parse has one scalar dependency per seam -- where the previous
// shard's last token ends (squeeze.h:377-394: i += len) -- which is resolved without a
// collective: every device composes its shard's exit map (overshoot for every possible entry,
// 1 KiB), the host chains the maps (one lookup per seam), every device then emits its tokens from
// its true entry, and the token arrays are concatenated in shard order by plain copies sized by
// their counts: device -> host when tokens_out is host memory, peer-to-peer (cudaMemcpyPeerAsync
// over NVLink) when it is device memory.
// ---------------------------------------------------------------------------
struct MultiPart {
    int device = 0;
    size_t first = 0, n = 0, back = 0, ahead = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t mapped = nullptr, parsed = nullptr;
    uint8_t* d_data = nullptr;
    uint32_t* d_table = nullptr;
    uint32_t* d_tokens = nullptr;
    void* d_work = nullptr;
    void* d_mwork = nullptr;
    uint16_t* d_map = nullptr;
    uint64_t* d_result = nullptr;
    uint16_t* h_map = nullptr;        // pinned: exit map
    uint64_t* h_result = nullptr;     // pinned: count, overshoot
    size_t cap_n = 0, cap_data = 0;   // what the buffers hold (parts are recycled between calls)
};

static void multi_free(MultiPart& p) {
    cudaSetDevice(p.device);
    if (p.stream) { cudaStreamSynchronize(p.stream); }
    cudaFree(p.d_data); cudaFree(p.d_table); cudaFree(p.d_tokens); cudaFree(p.d_work);
    cudaFree(p.d_mwork); cudaFree(p.d_map); cudaFree(p.d_result);
    cudaFreeHost(p.h_map); cudaFreeHost(p.h_result);
    if (p.mapped) { cudaEventDestroy(p.mapped); }
    if (p.parsed) { cudaEventDestroy(p.parsed); }
    if (p.stream) { cudaStreamDestroy(p.stream); }
    p = MultiPart();
}

// Like the slots of the single-device pipeline, the per-device buffers of a multi-device call are
// kept for the next call of a similar size (allocating ten gigabytes takes longer than searching a
// small input); at most kMultiParked of them, and only parts below kMultiParkBytes of device memory.
static std::vector<MultiPart> g_multi_parked;      // guarded by g_park_mu
constexpr size_t kMultiParked = 8;
constexpr size_t kMultiParkBytes = (size_t)12 << 30;

static size_t multi_bytes(const MultiPart& p) { return p.cap_data + p.cap_n * 8 + parse::workspace(p.cap_n) + sqz_gpu_match_workspace(p.cap_n); }

static int multi_alloc(MultiPart& p) {
    const size_t need_data = p.back + p.n + p.ahead + 64;
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        for (size_t k = 0; k < g_multi_parked.size(); k++) {
            const MultiPart& c = g_multi_parked[k];
            if (c.device == p.device && c.cap_n >= p.n && c.cap_n <= 2 * p.n + ((size_t)1 << 20) && c.cap_data >= need_data) {
                MultiPart t = c;
                g_multi_parked.erase(g_multi_parked.begin() + (long)k);
                t.first = p.first; t.n = p.n; t.back = p.back; t.ahead = p.ahead;
                p = t;
                CU(cudaSetDevice(p.device));
                return 0;
            }
        }
    }
    CU(cudaSetDevice(p.device));
    p.cap_n = (std::max<size_t>(p.n, 1) + ((size_t)1 << 20) - 1) >> 20 << 20;
    p.cap_data = p.cap_n + (size_t)sqz_gpu_max_dist_limit + sqz_gpu_max_len_limit + 64;
    CU(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&p.mapped, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&p.parsed, cudaEventDisableTiming));
    CU(cudaMalloc(&p.d_data, p.cap_data));
    CU(cudaMalloc(&p.d_table, p.cap_n * 4));
    CU(cudaMalloc(&p.d_tokens, p.cap_n * 4 + 16));
    CU(cudaMalloc(&p.d_work, parse::workspace(p.cap_n)));
    CU(cudaMalloc(&p.d_mwork, sqz_gpu_match_workspace(p.cap_n)));
    CU(cudaMalloc(&p.d_map, parse::kMapStride * 2));
    CU(cudaMalloc(&p.d_result, 16));
    CU(cudaHostAlloc(&p.h_map, parse::kMapStride * 2, cudaHostAllocDefault));
    CU(cudaHostAlloc(&p.h_result, 16, cudaHostAllocDefault));
    return 0;
}

static void multi_give_back(MultiPart& p) {
    if (p.stream == nullptr || p.cap_n == 0 || multi_bytes(p) > kMultiParkBytes) { multi_free(p); return; }
    cudaSetDevice(p.device);
    cudaStreamSynchronize(p.stream);
    MultiPart victim;
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        g_multi_parked.push_back(p);
        if (g_multi_parked.size() > kMultiParked) { victim = g_multi_parked.front(); g_multi_parked.erase(g_multi_parked.begin()); }
    }
    p = MultiPart();
    if (victim.stream != nullptr) { multi_free(victim); }
}

static void multi_release_all() {
    std::vector<MultiPart> all;
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        all.swap(g_multi_parked);
    }
    for (MultiPart& p : all) { multi_free(p); }
}

static int multi_run(std::vector<MultiPart>& parts, const uint8_t* data, uint32_t min_len, uint32_t max_len,
                     uint32_t max_dist, uint32_t* tokens_out, size_t tokens_cap, bool out_on_device,
                     int out_device, size_t* n_tokens, size_t* shard_tokens) {
    // phase A on every device at once: upload, match table, composed exit map
    for (MultiPart& p : parts) {
        if (p.n == 0) { continue; }
        if (int r = multi_alloc(p)) { return r; }
        CU(cudaMemcpyAsync(p.d_data, data + p.first - p.back, p.back + p.n + p.ahead, cudaMemcpyHostToDevice, p.stream));
        const uint8_t* d_shard = p.d_data + p.back;
        if (int r = sqz_gpu_match_table_device_ws(d_shard, p.back, p.n, p.ahead, min_len, max_len, max_dist,
                                                  p.d_table, p.d_mwork, p.stream)) { return r; }
        if (int r = sqz_gpu_parse_exit_map_device(p.d_table, p.n, min_len, max_len, p.d_work, p.d_map, p.stream)) { return r; }
        CU(cudaMemcpyAsync(p.h_map, p.d_map, (size_t)max_len * 2, cudaMemcpyDeviceToHost, p.stream));
        CU(cudaEventRecord(p.mapped, p.stream));
    }
    // seams: shard g+1 is entered where shard g's last token ends
    uint32_t entry = 0;
    for (MultiPart& p : parts) {
        if (p.n == 0) { continue; }
        CU(cudaSetDevice(p.device));
        CU(cudaEventSynchronize(p.mapped));
        const uint8_t* d_shard = p.d_data + p.back;
        if (entry >= max_len) { return fail(EIO, "seam entry out of range"); }
        const uint32_t next_entry = p.h_map[entry];
        if (int r = parse_launch(d_shard, p.d_table, p.n, nullptr, entry, min_len, max_len, p.d_tokens, p.n,
                                 p.d_work, p.d_result, p.stream)) { return r; }
        CU(cudaMemcpyAsync(p.h_result, p.d_result, 16, cudaMemcpyDeviceToHost, p.stream));
        CU(cudaEventRecord(p.parsed, p.stream));
        entry = next_entry;
    }
    // concatenation in shard order, every copy sized by its shard's count
    size_t total = 0;
    for (size_t g = 0; g < parts.size(); g++) {
        MultiPart& p = parts[g];
        size_t c = 0;
        if (p.n != 0) {
            CU(cudaSetDevice(p.device));
            CU(cudaEventSynchronize(p.parsed));
            c = (size_t)p.h_result[0];
            if (c > p.n) { return fail(EIO, "parse produced more tokens than positions"); }
            if (tokens_out != nullptr && total < tokens_cap) {
                const size_t take = std::min(c, tokens_cap - total);
                if (out_on_device) {
                    CU(cudaMemcpyPeerAsync(tokens_out + total, out_device, p.d_tokens, p.device, take * 4, p.stream));
                } else {
                    CU(cudaMemcpyAsync(tokens_out + total, p.d_tokens, take * 4, cudaMemcpyDeviceToHost, p.stream));
                }
            }
        }
        if (shard_tokens != nullptr) { shard_tokens[g] = c; }
        total += c;
    }
    for (MultiPart& p : parts) {
        if (p.n == 0) { continue; }
        CU(cudaSetDevice(p.device));
        CU(cudaStreamSynchronize(p.stream));
    }
    *n_tokens = total;
    if (entry != 0) { return fail(EIO, "the parse does not end at the end of the input"); }
    return total > tokens_cap ? fail(E2BIG, "token buffer too small") : 0;
}

extern "C" int sqz_gpu_tokens_multi(const int* devices, int n_devices, const uint8_t* data, size_t bytes,
                                    uint32_t window, uint32_t min_len, uint32_t max_len, uint32_t max_dist,
                                    uint32_t* tokens_out, size_t tokens_cap, size_t* n_tokens,
                                    size_t* shard_tokens) {
    if (n_tokens == nullptr) { return fail(EINVAL, "null n_tokens"); }
    *n_tokens = 0;
    if (devices == nullptr || n_devices < 1 || n_devices > kMaxDevices) { return fail(EINVAL, "bad device list"); }
    if (int r = check_rules(min_len, max_len, max_dist)) { return r; }
    if (window == 0 || (window & (window - 1)) != 0 || max_dist > window) {
        return fail(EINVAL, "window must be a power of two and max_dist <= window");
    }
    if (data == nullptr && bytes != 0) { return fail(EINVAL, "null input"); }
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0) {
        return fail(ENODEV, "no CUDA device: the match search has no CPU fallback", ce);
    }
    for (int g = 0; g < n_devices; g++) {
        // (a device may be listed more than once: its shards then share it, each on its own stream)
        if (devices[g] < 0 || devices[g] >= count) { return fail(ENODEV, "no such CUDA device"); }
    }
    DeviceScope keep;
    // where do the tokens go: host memory, or device memory (then peer copies gather them there)
    bool out_on_device = false;
    int out_device = devices[0];
    if (tokens_out != nullptr) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, tokens_out) == cudaSuccess && at.type == cudaMemoryTypeDevice) {
            out_on_device = true;
            out_device = at.device;
        }
        cudaGetLastError();
    }
    if (out_on_device) {
        for (int g = 0; g < n_devices; g++) {
            if (devices[g] == out_device) { continue; }
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, devices[g], out_device));
            if (can) {
                CU(cudaSetDevice(devices[g]));
                ce = cudaDeviceEnablePeerAccess(out_device, 0);
                if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled) { return fail(cuda_code(ce), "cudaDeviceEnablePeerAccess", ce); }
                cudaGetLastError();
            }   // without peer access cudaMemcpyPeerAsync still works, staged through the host
        }
    }
    std::vector<MultiPart> parts((size_t)n_devices);
    const size_t base = bytes / (size_t)n_devices, extra = bytes % (size_t)n_devices;
    size_t first = 0;
    for (int g = 0; g < n_devices; g++) {
        MultiPart& p = parts[(size_t)g];
        p.device = devices[g];
        p.first = first;
        p.n = base + ((size_t)g < extra ? 1 : 0);
        p.back = std::min<size_t>(first, max_dist);
        p.ahead = std::min<size_t>(bytes - (first + p.n), max_len);
        first += p.n;
    }
    const int rc = multi_run(parts, data, min_len, max_len, max_dist, tokens_out, tokens_cap, out_on_device,
                             out_device, n_tokens, shard_tokens);
    for (MultiPart& p : parts) {
        if (rc == 0 && p.stream != nullptr) { multi_give_back(p); }
        else if (p.stream != nullptr || p.d_data != nullptr) { multi_free(p); }
    }
    return rc;
}

// ---------------------------------------------------------------------------
// one process per device (torchrun-style jobs): the same gather needs the destination buffer of
// one process mapped into the others.  Plain CUDA IPC, no collective library.
// ---------------------------------------------------------------------------
extern "C" int sqz_gpu_device_alloc(void** d_ptr, size_t bytes) {
    if (d_ptr == nullptr) { return fail(EINVAL, "null argument"); }
    *d_ptr = nullptr;
    CU(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return 0;
}

extern "C" void sqz_gpu_device_free(void* d_ptr) {
    if (d_ptr != nullptr) { cudaFree(d_ptr); }
}

extern "C" int sqz_gpu_ipc_export(const void* d_ptr, uint8_t handle[SQZ_GPU_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == SQZ_GPU_IPC_HANDLE_BYTES, "IPC handle size");
    if (d_ptr == nullptr || handle == nullptr) { return fail(EINVAL, "null argument"); }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(handle, &h, sizeof(h));
    return 0;
}

extern "C" int sqz_gpu_ipc_open(const uint8_t handle[SQZ_GPU_IPC_HANDLE_BYTES], void** d_ptr) {
    if (d_ptr == nullptr || handle == nullptr) { return fail(EINVAL, "null argument"); }
    *d_ptr = nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int sqz_gpu_ipc_close(void* d_ptr) {
    if (d_ptr == nullptr) { return 0; }
    CU(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}

// copy `count` tokens of this process's device into a (possibly IPC-mapped, possibly remote)
// device buffer at token offset `at`: the per-shard step of the gather
extern "C" int sqz_gpu_put_tokens(uint32_t* d_dst, size_t at, const uint32_t* d_tokens, size_t count, void* stream) {
    if (count == 0) { return 0; }
    if (d_dst == nullptr || d_tokens == nullptr) { return fail(EINVAL, "null argument"); }
    CU(cudaMemcpyAsync(d_dst + at, d_tokens, count * 4, cudaMemcpyDefault, (cudaStream_t)stream));
    return 0;
}

// ---------------------------------------------------------------------------
// decoder side: the LZ copy phase on the GPU (lz_expand.cuh)
// ---------------------------------------------------------------------------
static size_t expand_blocks(size_t n_tokens) { return (n_tokens + expand::kScanBlock - 1) / expand::kScanBlock; }

extern "C" size_t sqz_gpu_expand_workspace(size_t n_tokens, size_t bytes) {
    // block offsets + {total, bad, changed} + one 32-bit hop per output byte
    return parse::align_up(expand_blocks(n_tokens) * 8 + 8) + 256 + parse::align_up(bytes * 4 + 4);
}

extern "C" int sqz_gpu_expand_tokens_device(const uint32_t* d_tokens, size_t n_tokens, uint8_t* d_out,
                                            size_t bytes, void* d_work, void* stream) {
    if (bytes >= ((size_t)1 << 31)) { return fail(EINVAL, "one expand call handles less than 2 GiB"); }
    if (n_tokens > bytes) { return fail(EINVAL, "more tokens than bytes"); }
    if (bytes == 0) { return 0; }
    if (n_tokens == 0) { return fail(EINVAL, "tokens describe fewer bytes than announced"); }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t blocks = expand_blocks(n_tokens);
    if (blocks > 0x7FFFFFFFull) { return fail(EINVAL, "too many tokens for one launch"); }
    uint8_t* w = static_cast<uint8_t*>(d_work);
    uint64_t* block_off = reinterpret_cast<uint64_t*>(w);
    uint8_t* flags = w + parse::align_up(blocks * 8 + 8);
    uint64_t* d_total = reinterpret_cast<uint64_t*>(flags);
    int* d_bad = reinterpret_cast<int*>(flags + 8);
    int* d_rounds = reinterpret_cast<int*>(flags + 16);          // three flags of the doubling rounds
    uint32_t* hop = reinterpret_cast<uint32_t*>(flags + 256);
    CU(cudaMemsetAsync(flags, 0, 256, s));
    expand::block_lengths<<<(unsigned)blocks, 256, 0, s>>>(d_tokens, n_tokens, block_off);
    LAUNCHED("expand_block_lengths");
    expand::scan_blocks<<<1, 1024, 0, s>>>(block_off, blocks, d_total);
    LAUNCHED("expand_scan_blocks");
    struct { uint64_t total; int bad; int pad; int changed; } h = { 0, 0, 0, 0 };
    CU(cudaMemcpyAsync(&h, flags, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (h.total != bytes) { return fail(EINVAL, "tokens do not describe the announced number of bytes"); }
    expand::place_tokens<<<(unsigned)blocks, 256, 0, s>>>(d_tokens, n_tokens, block_off, d_out, hop,
                                                        (uint64_t)bytes, d_bad);
    LAUNCHED("expand_place_tokens");
    const unsigned grid = (unsigned)sm_count() * 8;
    // a chain halves its hop count every round: 32 rounds cover 31-bit hops; rounds after the last
    // useful one return at once (the device decides, the host does not wait in between)
    for (int round = 0; round < 32; round++) {
        expand::double_hops<<<grid, 256, 0, s>>>(hop, (uint64_t)bytes, d_rounds, round);
        LAUNCHED("expand_double_hops");
    }
    expand::fetch_bytes<<<grid, 256, 0, s>>>(hop, d_out, (uint64_t)bytes);
    LAUNCHED("expand_fetch_bytes");
    // one wait for the whole copy phase: was a token out of range?  (The bytes of such a stream are
    // garbage but in bounds: place_tokens cut every bad hop to 0.)
    CU(cudaMemcpyAsync(&h, flags, 16, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (h.bad != 0) { return fail(EINVAL, "a match reaches before the start of the output"); }
    return 0;
}

extern "C" int sqz_gpu_expand_tokens(const uint32_t* tokens, size_t n_tokens, uint8_t* out, size_t bytes) {
    if (bytes == 0) { return n_tokens == 0 ? 0 : fail(EINVAL, "tokens for an empty output"); }
    if (tokens == nullptr || out == nullptr) { return fail(EINVAL, "null argument"); }
    int device = -1;
    if (int r = ensure_device(device)) { return r; }
    uint32_t* d_tokens = nullptr;
    uint8_t* d_out = nullptr;
    void* d_work = nullptr;
    cudaStream_t s = nullptr;
    int rc = 0;
    cudaError_t ce = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (ce == cudaSuccess) { ce = cudaMalloc(&d_tokens, n_tokens * 4 + 4); }
    if (ce == cudaSuccess) { ce = cudaMalloc(&d_out, bytes); }
    if (ce == cudaSuccess) { ce = cudaMalloc(&d_work, sqz_gpu_expand_workspace(n_tokens, bytes)); }
    if (ce == cudaSuccess) { ce = cudaMemcpyAsync(d_tokens, tokens, n_tokens * 4, cudaMemcpyHostToDevice, s); }
    if (ce != cudaSuccess) { rc = fail(cuda_code(ce), "expand: allocation or upload", ce); }
    if (rc == 0) { rc = sqz_gpu_expand_tokens_device(d_tokens, n_tokens, d_out, bytes, d_work, s); }
    if (rc == 0) {
        ce = cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess) { ce = cudaStreamSynchronize(s); }
        if (ce != cudaSuccess) { rc = fail(cuda_code(ce), "expand: download", ce); }
    }
    cudaFree(d_tokens); cudaFree(d_out); cudaFree(d_work);
    if (s != nullptr) { cudaStreamDestroy(s); }
    return rc;
}

// ---------------------------------------------------------------------------
// utilities
// ---------------------------------------------------------------------------
extern "C" int sqz_gpu_abi_version(void) { return SQZ_GPU_ABI_VERSION; }

extern "C" int sqz_gpu_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

extern "C" const char* sqz_gpu_last_error(void) { return g_err; }

extern "C" void* sqz_gpu_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void sqz_gpu_host_free(void* p) {
    if (p != nullptr) { cudaFreeHost(p); }
}

extern "C" uint64_t sqz_gpu_launch_count(void) { return g_launches.load(); }
