/* oracle/sqz_oracle.c -- TEST INFRASTRUCTURE (the checker), never shipped.
 *
 * CPU restatement of the one hot path of leok7v/sqz that the CUDA build
 * replaces.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may call it; the product (libsqz_b200.so) never links or loads it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *  (a) the unmodified reference compiled into oracle/_ref (ref_tokens: the
 *      reference's own decision at every parse position of every fixture and
 *      of shifted suffix slices of them; ref_compress byte equality of
 *      "oracle tokens fed through the reference's own encoder"),
 *  (b) the reference's second brute-force loop, bst.c:230-252, for rule set iii,
 *  (c) the committed golden digests under tests/golden/ (made here with
 *      oracle/_ref by tests/golden/make_golden.py).
 *
 * Reference semantics restated (attic/map_experiment/squeeze.h):
 *   :340      position 0 never searches (literal)
 *   :341-342  candidates run from j = i-1 (nearest) down to
 *             min_j = i >= window ? i-window+1 : 0, i.e. dist in [1, window-1]
 *   :345-349  k = equal bytes of data[j..] and data[i..], capped by bytes-i and
 *             by max_len; the compare may run past i (overlap is allowed)
 *   :350-354  a candidate wins only if k >= min_len and k > best so far, so
 *             among equal lengths the nearest stays; stop once best == max_len
 *   :377-394  greedy: len >= min_len emits (len,dist) and skips len bytes,
 *             otherwise one literal
 */
#include "sqz_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define ORACLE_API __attribute__((visibility("default")))

static inline uint32_t common_prefix(const uint8_t* a, const uint8_t* b,
                                     uint32_t limit) {
    uint32_t k = 0;
    while (k < limit && a[k] == b[k]) { k++; }
    return k;
}

ORACLE_API void oracle_best(const uint8_t* data, uint64_t bytes, uint64_t i,
                            const oracle_rules* r, uint32_t* len, uint32_t* dist) {
    uint32_t best = 0, where = 0;
    if (i > 0 && i < bytes) {
        const uint64_t room = bytes - i;
        const uint32_t limit = room < r->max_len ? (uint32_t)room : r->max_len;
        const uint64_t reach = i < r->max_dist ? i : r->max_dist;
        for (uint64_t d = 1; d <= reach; d++) {          /* nearest first */
            uint32_t k = common_prefix(data + i - d, data + i, limit);
            if (k >= r->min_len && k > best) {            /* strictly longer */
                best = k; where = (uint32_t)d;
                if (best == r->max_len) { break; }
            }
        }
    }
    *len = best; *dist = where;
}

/* ---- tiny pthread parallel-for (this image has no libgomp) ------------------ */
typedef void (*range_fn)(void* ctx, uint64_t lo, uint64_t hi);
typedef struct {
    range_fn fn; void* ctx; uint64_t count, grain; uint64_t* next;
    pthread_mutex_t* mu;
} pf_job;

static void* pf_worker(void* arg) {
    pf_job* j = (pf_job*)arg;
    for (;;) {
        pthread_mutex_lock(j->mu);
        uint64_t lo = *j->next;
        *j->next = lo + j->grain;
        pthread_mutex_unlock(j->mu);
        if (lo >= j->count) { break; }
        uint64_t hi = lo + j->grain < j->count ? lo + j->grain : j->count;
        j->fn(j->ctx, lo, hi);
    }
    return NULL;
}

static int oracle_threads = 0; /* 0 = all online cores */

ORACLE_API void oracle_set_threads(int n) { oracle_threads = n; }

ORACLE_API int oracle_get_threads(void) {
    if (oracle_threads > 0) { return oracle_threads; }
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void parallel_for(uint64_t count, uint64_t grain, range_fn fn, void* ctx) {
    int nt = oracle_get_threads();
    if (nt > 64) { nt = 64; }
    if (count <= grain || nt <= 1) { fn(ctx, 0, count); return; }
    pthread_t th[64];
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    uint64_t next = 0;
    pf_job job = { fn, ctx, count, grain, &next, &mu };
    int started = 0;
    for (int t = 0; t < nt - 1; t++) {
        if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) { started++; }
    }
    pf_worker(&job);
    for (int t = 0; t < started; t++) { pthread_join(th[t], NULL); }
}

typedef struct {
    const uint8_t* data; uint64_t bytes; const oracle_rules* r; uint64_t first;
    uint16_t* len_out; uint16_t* dist_out;
    /* oracle B only */
    const uint64_t* prev; uint64_t lo;
} table_ctx;

static void brute_range(void* vc, uint64_t lo, uint64_t hi) {
    table_ctx* c = (table_ctx*)vc;
    for (uint64_t k = lo; k < hi; k++) {
        uint32_t l, d;
        oracle_best(c->data, c->bytes, c->first + k, c->r, &l, &d);
        c->len_out[k] = (uint16_t)l; c->dist_out[k] = (uint16_t)d;
    }
}

ORACLE_API void oracle_match_table(const uint8_t* data, uint64_t bytes,
                                   const oracle_rules* r, uint64_t first,
                                   uint64_t count, uint16_t* len_out,
                                   uint16_t* dist_out) {
    table_ctx c = { data, bytes, r, first, len_out, dist_out, NULL, 0 };
    parallel_for(count, 256, brute_range, &c);
}

ORACLE_API uint64_t oracle_tokens(const uint8_t* data, uint64_t bytes,
                                  const oracle_rules* r, uint32_t* tokens,
                                  uint64_t cap) {
    uint64_t i = 0, n = 0;
    while (i < bytes) {
        uint32_t l, d;
        oracle_best(data, bytes, i, r, &l, &d);
        uint32_t t;
        if (l >= r->min_len) { t = (l << 16) | d; i += l; }
        else                 { t = data[i];       i += 1; }
        if (n < cap) { tokens[n] = t; }
        n++;
    }
    return n;
}

ORACLE_API uint64_t oracle_tokens_from_table(const uint8_t* data, uint64_t bytes,
                                             const uint16_t* len, const uint16_t* dist,
                                             uint32_t min_len, uint64_t start,
                                             uint32_t* tokens, uint64_t cap,
                                             uint64_t* end_pos) {
    uint64_t i = start, n = 0;
    while (i < bytes) {
        uint32_t t;
        if (len[i] >= min_len) { t = ((uint32_t)len[i] << 16) | dist[i]; i += len[i]; }
        else                   { t = data[i];                             i += 1; }
        if (n < cap) { tokens[n] = t; }
        n++;
    }
    if (end_pos) { *end_pos = i; }
    return n;
}

/* ---- oracle B: exact hash chains -------------------------------------------
 * head[h] / prev[p] link every earlier position with the same min_len-byte
 * prefix hash, newest first, with NO chain-length limit, so walking a chain
 * visits exactly the candidates that can reach min_len, nearest first.  The
 * acceptance rule is the same strict '>' so the result equals oracle_best.
 * Chains are built for the whole buffer once; a query at i only follows links
 * to positions < i within max_dist.                                          */
static inline uint32_t hash_prefix(const uint8_t* p, uint32_t n) {
    uint32_t h = 2166136261u;
    for (uint32_t k = 0; k < n; k++) { h = (h ^ p[k]) * 16777619u; }
    return h >> 10; /* 22 bits */
}

static void chain_range(void* vc, uint64_t from, uint64_t to) {
    table_ctx* c = (table_ctx*)vc;
    const uint8_t* data = c->data; const oracle_rules* r = c->r;
    const uint64_t bytes = c->bytes, lo = c->lo, NIL = ~(uint64_t)0;
    const uint64_t* prev = c->prev;
    for (uint64_t k = from; k < to; k++) {
        const uint64_t i = c->first + k;
        uint32_t best = 0, where = 0;
        if (i > 0 && i < bytes && bytes - i >= r->min_len) {
            const uint64_t room = bytes - i;
            const uint32_t limit = room < r->max_len ? (uint32_t)room : r->max_len;
            uint64_t j = prev[i - lo];
            while (j != NIL && i - j <= r->max_dist) {
                /* cheap reject: to beat 'best' the byte at offset best must match */
                if (best < limit && data[j + best] == data[i + best]) {
                    uint32_t m = common_prefix(data + j, data + i, limit);
                    if (m >= r->min_len && m > best) {
                        best = m; where = (uint32_t)(i - j);
                        if (best == r->max_len) { break; }
                    }
                }
                if (best >= limit) { break; }
                j = prev[j - lo];
            }
        }
        c->len_out[k] = (uint16_t)best; c->dist_out[k] = (uint16_t)where;
    }
}

ORACLE_API void oracle_fast_table(const uint8_t* data, uint64_t bytes,
                                  const oracle_rules* r, uint64_t first,
                                  uint64_t count, uint16_t* len_out,
                                  uint16_t* dist_out) {
    const uint32_t HB = 1u << 22;
    const uint64_t NIL = ~(uint64_t)0;
    /* build links over [lo, first+count): lo far enough back for every query */
    uint64_t lo = first > r->max_dist ? first - r->max_dist : 0;
    uint64_t hi = first + count;
    if (hi > bytes) { hi = bytes; }
    uint64_t span = hi > lo ? hi - lo : 0;
    uint64_t* head = (uint64_t*)malloc(sizeof(uint64_t) * HB);
    uint64_t* prev = (uint64_t*)malloc(sizeof(uint64_t) * (span ? span : 1));
    if (!head || !prev) { abort(); }
    for (uint32_t k = 0; k < HB; k++) { head[k] = NIL; }
    for (uint64_t p = lo; p < hi; p++) {
        if (bytes - p >= r->min_len) {
            uint32_t h = hash_prefix(data + p, r->min_len);
            prev[p - lo] = head[h];
            head[h] = p;
        } else {
            prev[p - lo] = NIL;
        }
    }
    table_ctx c = { data, bytes, r, first, len_out, dist_out, prev, lo };
    parallel_for(count, 4096, chain_range, &c);
    free(head); free(prev);
}

ORACLE_API uint64_t oracle_fnv1a64(const void* p, uint64_t bytes) {
    const uint8_t* b = (const uint8_t*)p;
    uint64_t h = 0xCBF29CE484222325ull;
    for (uint64_t k = 0; k < bytes; k++) { h = (h ^ b[k]) * 0x100000001B3ull; }
    return h;
}
