/* sqz_codec.c -- host half of sqz-b200: header, bit I/O, the adaptive-Huffman
 * token coder and the decompressor.  Written from the behavioural description
 * in SURVEY.md Appendix A; every emitted bit has to equal what the reference
 * (/root/reference/attic/map_experiment/{squeeze,huffman,bitstream}.h) emits,
 * which tests/test_codec.py checks against oracle/_ref and tests/golden.
 *
 * The LZ77 search that feeds sqz_compress() lives on the GPU (sqz_gpu.cu).
 * This file never searches: no CPU fallback exists for that step.
 */
#include "sqz.h"
#include "sqz_gpu.h"

#include <errno.h>
#include <string.h>
#include <time.h>

/* ---- deflate-style bucket tables (RFC 1951 3.2.5; reference squeeze.h:29-79) */
static const uint16_t len_base[29] = {
    3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
    67, 83, 99, 115, 131, 163, 195, 227, 258 };
static const uint8_t len_extra[29] = {
    0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3,
    4, 4, 4, 4, 5, 5, 5, 5, 0 };
static const uint16_t pos_base[30] = {
    1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513,
    769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
static const uint8_t pos_extra[30] = {
    0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
    9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };

enum { len_symbol0 = 257 };

static double now_seconds(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ======================================================================== *
 *  bit I/O  (reference bitstream.h:28-114)                                  *
 *  Bits enter b64 from the right, so the first bit written ends up as the   *
 *  MSB of a 64-bit word; values are fed least-significant bit first.        *
 * ======================================================================== */

static inline uint64_t reverse64(uint64_t v) {
    v = ((v >> 1)  & 0x5555555555555555ull) | ((v & 0x5555555555555555ull) << 1);
    v = ((v >> 2)  & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
    v = ((v >> 4)  & 0x0F0F0F0F0F0F0F0Full) | ((v & 0x0F0F0F0F0F0F0F0Full) << 4);
    return __builtin_bswap64(v);
}

static void word_out(struct sqz_bitstream* bs) {
    if (bs->data != NULL && bs->capacity > 0) {
        for (int k = 7; k >= 0 && bs->error == 0; k--) {      /* big-endian */
            if (bs->bytes == bs->capacity) { bs->error = E2BIG; }
            else { bs->data[bs->bytes++] = (uint8_t)(bs->b64 >> (k * 8)); }
        }
    } else if (bs->output != NULL) {
        bs->error = bs->output(bs);                           /* host order */
        if (bs->error == 0) { bs->bytes += 8; }
    } else {
        bs->error = EINVAL;
    }
    bs->bits = 0;
    bs->b64 = 0;
}

/* append the low `count` bits of `value`, least significant first */
static inline void put_bits(struct sqz_bitstream* bs, uint64_t value, int count) {
    if (bs->error != 0 || count <= 0) { return; }
    /* in emission order the value is bit-reversed: its bit 0 goes out first */
    uint64_t seq = reverse64(value) >> (64 - count);
    int room = 64 - bs->bits;
    if (count < room) {
        bs->b64 = (bs->b64 << count) | seq;
        bs->bits += count;
        return;
    }
    int rest = count - room;                                  /* bits left over */
    bs->b64 = (room == 64 ? 0 : bs->b64 << room) | (seq >> rest);
    word_out(bs);
    if (rest > 0 && bs->error == 0) {
        bs->b64 = seq & (((uint64_t)1 << rest) - 1);
        bs->bits = rest;
    }
}

static inline void pad_to_word(struct sqz_bitstream* bs) {     /* bitstream.h:112-114 */
    if (bs->error == 0 && bs->bits > 0) { put_bits(bs, 0, 64 - bs->bits); }
}

static void word_in(struct sqz_bitstream* bs) {
    bs->b64 = 0;
    if (bs->data != NULL && bs->bytes > 0) {
        for (int k = 7; k >= 0 && bs->error == 0; k--) {
            if (bs->read == bs->bytes) { bs->error = E2BIG; }
            else { bs->b64 |= (uint64_t)bs->data[bs->read++] << (k * 8); }
        }
    } else if (bs->input != NULL) {
        bs->error = bs->input(bs);
        if (bs->error == 0) { bs->read += 8; }
    } else {
        bs->error = EINVAL;
    }
    bs->bits = 64;
}

static inline int get_bit(struct sqz_bitstream* bs) {
    if (bs->error != 0) { return 0; }
    if (bs->bits == 0) { word_in(bs); if (bs->error != 0) { return 0; } }
    int bit = (int)(bs->b64 >> 63);
    bs->b64 <<= 1;
    bs->bits--;
    return bit;
}

static inline uint64_t get_bits(struct sqz_bitstream* bs, int count) {
    uint64_t v = 0;
    if (bs->error == 0 && bs->bits >= count && count > 0 && count < 64) {
        /* whole field is in the register: top `count` bits, first bit = LSB */
        uint64_t top = bs->b64 >> (64 - count);
        bs->b64 <<= count;
        bs->bits -= count;
        return reverse64(top) >> (64 - count);
    }
    for (int k = 0; k < count && bs->error == 0; k++) {
        v |= (uint64_t)get_bit(bs) << k;
    }
    return v;
}

/* ======================================================================== *
 *  adaptive Huffman tree  (reference huffman.h:36-269)    *
 *  Leaves are 0..n-1 (index == symbol); the root is 2n-2; further internal  *
 *  nodes are allocated downward from 2n-3.                                  *
 * ======================================================================== */

static void tree_init(struct sqz_tree* t, struct sqz_node* nodes, int32_t leaves) {
    t->node = nodes;
    t->n = leaves;
    t->next = 2 * leaves - 2;
    t->depth = 0;
    t->complete = 0;
    for (int32_t k = 0; k < 2 * leaves - 1; k++) {
        nodes[k].freq = 0; nodes[k].path = 0; nodes[k].bits = 0;
        nodes[k].up = -1;  nodes[k].lo = -1;  nodes[k].hi = -1;
    }
}

static inline int32_t tree_root(const struct sqz_tree* t) { return 2 * t->n - 2; }

/* Re-derive code length and code of everything below `top` from top's own.
 * A relabel that starts at the root restarts the depth high-water mark
 * (huffman.h:41-62).  Iterative: the order of visits does not matter.       */
static void relabel(struct sqz_tree* t, int32_t top) {
    struct sqz_node* nd = t->node;
    int32_t stack[2 * sqz_lit_symbols];
    int sp = 0;
    int32_t depth = (top == tree_root(t)) ? 0 : t->depth;
    stack[sp++] = top;
    while (sp > 0) {
        const int32_t i = stack[--sp];
        const int32_t bits = nd[i].bits;
        const uint64_t path = nd[i].path;
        if (bits > depth) { depth = bits; }
        const int32_t lo = nd[i].lo, hi = nd[i].hi;
        if (lo >= 0) { nd[lo].bits = bits + 1; nd[lo].path = path; stack[sp++] = lo; }
        if (hi >= 0) { nd[hi].bits = bits + 1; nd[hi].path = path | ((uint64_t)1 << bits); stack[sp++] = hi; }
    }
    t->depth = depth;
}

static inline void sum_children(struct sqz_tree* t, int32_t i) {
    const struct sqz_node* nd = t->node;
    t->node[i].freq = (nd[i].lo >= 0 ? nd[nd[i].lo].freq : 0) +
                      (nd[i].hi >= 0 ? nd[nd[i].hi].freq : 0);
}

/* Keep the lighter child on the left.  When the children trade places the
 * caller continues with the node that now sits where `i` used to be, i.e.
 * i's sibling (huffman.h:64-86).                                            */
static int32_t order_siblings(struct sqz_tree* t, int32_t i) {
    struct sqz_node* nd = t->node;
    if (i == tree_root(t)) { return i; }
    const int32_t p = nd[i].up;
    const int32_t lo = nd[p].lo, hi = nd[p].hi;
    if (lo >= 0 && hi >= 0 && nd[lo].freq > nd[hi].freq) {
        nd[p].lo = hi;
        nd[p].hi = lo;
        relabel(t, p);
        return i == lo ? hi : lo;
    }
    return i;
}

static void weight_changed(struct sqz_tree* t, int32_t i);

/* `x` is the right child of p; if it outweighs p's sibling u the two trade
 * places: x moves up next to p, u moves down under p (huffman.h:98-128).    */
static void promote(struct sqz_tree* t, int32_t x) {
    struct sqz_node* nd = t->node;
    const int32_t p = nd[x].up;
    const int32_t g = nd[p].up;
    const int p_left = nd[g].lo == p;
    const int32_t u = p_left ? nd[g].hi : nd[g].lo;
    if (nd[x].freq > nd[u].freq) {
        nd[x].up = g;
        if (p_left) { nd[g].hi = x; } else { nd[g].lo = x; }
        nd[p].hi = u;
        nd[u].up = p;
        sum_children(t, p);
        sum_children(t, g);
        (void)order_siblings(t, x);
        (void)order_siblings(t, u);
        (void)order_siblings(t, p);
        relabel(t, g);
        weight_changed(t, g);
    }
}

/* Propagate a weight change from `i` to the root, re-ordering siblings on the
 * way up and, on the way back down, promoting right children that outgrew
 * their uncle (huffman.h:130-147).  The reference recurses; this is the same
 * sequence of steps with the recursion unrolled: first every level from the
 * leaf to the root refreshes its parent's weight and orders the two children
 * (continuing, after a swap, with the node that took the old slot), then the
 * levels are revisited from the root down for the promotion test, each with
 * the (node, parent) pair it captured on the way up.
 *
 * The second pass is almost always a no-op, and whether it is can be told on
 * the way up at no cost: the test of level k reads only the final weight of
 * its node (settled at level k), whether that node ended up as the right
 * child (settled at level k) and the weight of the parent's sibling, which
 * is not on the path and is loaded anyway at level k+1.  Nothing changes the
 * tree during the second pass unless a test fires, so if no test would fire
 * on the state left by the first pass the second pass is skipped; otherwise it
 * runs exactly as the reference's unwinding does.                            */
static void weight_changed(struct sqz_tree* t, int32_t i) {
    struct sqz_node* nd = t->node;
    int32_t node_at[2 * sqz_lit_symbols], parent_at[2 * sqz_lit_symbols];
    int levels = 0;
    int may_promote = 0;
    int below_is_right = 0;             /* level below: did its node end up as the right child? */
    uint64_t below_weight = 0;          /* level below: final weight of its node */
    for (;;) {
        const int32_t p = nd[i].up;
        if (p < 0) {                    /* the root: refresh its own weight, nothing to order */
            sum_children(t, i);
            break;
        }
        const int32_t lo = nd[p].lo, hi = nd[p].hi;
        const uint64_t wl = lo >= 0 ? nd[lo].freq : 0, wh = hi >= 0 ? nd[hi].freq : 0;
        /* `i` (the parent of the level below) has a sibling here: the promotion test of the
         * level below compares against its weight */
        if (below_is_right && below_weight > (i == lo ? wh : wl)) { may_promote = 1; }
        nd[p].freq = wl + wh;
        int32_t right = hi;
        if (lo >= 0 && hi >= 0 && wl > wh) { /* heavier child goes right */
            nd[p].lo = hi;
            nd[p].hi = lo;
            relabel(t, p);
            i = (i == lo) ? hi : lo;
            right = lo;
        }
        below_is_right = (i == right);
        below_weight = (i == lo) ? wl : wh;
        node_at[levels] = i;
        parent_at[levels] = p;
        levels++;
        i = p;
    }
    if (!may_promote) { return; }
    while (levels > 0) {
        levels--;
        const int32_t p = parent_at[levels];
        if (nd[p].up >= 0 && nd[p].hi == node_at[levels]) { promote(t, node_at[levels]); }
    }
}

/* First occurrence of symbol `s` (huffman.h:149-216): walk from the root,
 * always to the left, to the first free child slot (right slot preferred) or
 * to a leaf, which is then split by a fresh internal node.                  */
static int tree_insert(struct sqz_tree* t, int32_t s) {
    struct sqz_node* nd = t->node;
    int ok = 1;
    int32_t at = tree_root(t);
    nd[s].freq = 1;
    while (at >= t->n) {
        if (nd[at].hi < 0)      { nd[at].hi = s; nd[s].up = at; break; }
        else if (nd[at].lo < 0) { nd[at].lo = s; nd[s].up = at; break; }
        else                    { at = nd[at].lo; }
    }
    if (at >= t->n) {
        nd[at].freq++;
        s = order_siblings(t, s);
    } else if (t->next == t->n) {
        ok = 0;
        t->complete = 1;
    } else {
        const int32_t leaf = at;
        const int32_t x = --t->next;
        nd[x].freq = nd[leaf].freq;
        nd[x].path = nd[leaf].path;
        nd[x].bits = nd[leaf].bits;
        nd[x].up   = nd[leaf].up;
        nd[x].lo   = leaf;
        nd[x].hi   = s;
        if (nd[x].up >= 0) {
            if (nd[nd[x].up].lo == leaf) { nd[nd[x].up].lo = x; } else { nd[nd[x].up].hi = x; }
        }
        nd[leaf].up = x;
        nd[leaf].bits = nd[x].bits + 1;          /* left edge: same code, one longer */
        nd[s].up = x;
        nd[s].bits = nd[x].bits + 1;
        nd[s].path = nd[x].path | ((uint64_t)1 << nd[x].bits);
        sum_children(t, x);
        at = x;
    }
    weight_changed(t, s);
    relabel(t, at);
    return ok;
}

static void tree_count(struct sqz_tree* t, int32_t s) {        /* huffman.h:218-235 */
    struct sqz_node* nd = t->node;
    if (nd[s].up < 0) {
        (void)tree_insert(t, s);
    } else if (!t->complete && t->depth < 63 && nd[s].freq < UINT64_MAX - 1) {
        nd[s].freq++;
        weight_changed(t, s);
    } else {
        t->complete = 1;
    }
}

/* ======================================================================== *
 *  token coder  (reference squeeze.h:151-172, 239-315)                      *
 * ======================================================================== */

static void bucket_tables(struct sqz* s) {
    /* len -> bucket: 258 would be bucket 28 in deflate, but symbol 257+28 is
     * the NYT escape here, so 227..258 all stay in bucket 27 (squeeze.h:151-161) */
    memset(s->len_index, 0, sizeof(s->len_index));
    for (int len = 3; len < (int)sizeof(s->len_index); len++) {
        int b = 0;
        while (b + 1 < 28 && len_base[b + 1] <= len) { b++; }
        s->len_index[len] = (uint8_t)b;
    }
    int b = 0;
    for (uint32_t d = 0; d < (1u << 15); d++) {
        while (b + 1 < 30 && pos_base[b + 1] <= d) { b++; }
        s->pos_index[d] = (uint8_t)b;
    }
}

void sqz_init(struct sqz* s) {
    memset(s, 0, sizeof(*s));
    s->device = -1;                       /* the caller's current CUDA device */
    tree_init(&s->lit, s->lit_nodes, sqz_lit_symbols);
    tree_init(&s->pos, s->pos_nodes, sqz_pos_symbols);
}

static void coder_begin(struct sqz* s, struct sqz_bitstream* bs) {
    s->bs = bs;
    /* both escape symbols exist from the start, so the first code ever
     * written is the single bit 1 (squeeze.h:333-334) */
    if (!tree_insert(&s->lit, sqz_lit_nyt)) { s->error = EINVAL; }
    if (!tree_insert(&s->pos, sqz_pos_nyt)) { s->error = EINVAL; }
    bucket_tables(s);
}

static inline void s_put(struct sqz* s, uint64_t v, int count) {
    if (s->error == 0) { put_bits(s->bs, v, count); s->error = s->bs->error; }
}

/* current code of `sym`, then bump its weight (squeeze.h:239-246) */
static inline void emit_symbol(struct sqz* s, struct sqz_tree* t, int32_t sym) {
    s_put(s, t->node[sym].path, t->node[sym].bits);
    tree_count(t, sym);
}

static inline void code_lit(struct sqz* s, uint32_t sym) {     /* squeeze.h:278-288 */
    if (s->lit.node[sym].bits == 0) {
        emit_symbol(s, &s->lit, sqz_lit_nyt);
        s_put(s, sym, 9);
        if (!tree_insert(&s->lit, (int32_t)sym)) { s->error = E2BIG; }
    } else {
        emit_symbol(s, &s->lit, (int32_t)sym);
    }
}

static inline void code_len(struct sqz* s, uint32_t len) {     /* squeeze.h:290-298 */
    const uint32_t b = s->len_index[len];
    code_lit(s, len_symbol0 + b);
    if (len_extra[b] > 0) { s_put(s, len - len_base[b], len_extra[b]); }
}

static inline void code_dist(struct sqz* s, uint32_t dist) {   /* squeeze.h:300-315 */
    const uint32_t b = s->pos_index[dist];
    if (s->pos.node[b].bits == 0) {
        emit_symbol(s, &s->pos, sqz_pos_nyt);
        s_put(s, b, 5);
        if (!tree_insert(&s->pos, (int32_t)b)) { s->error = E2BIG; }
    } else {
        emit_symbol(s, &s->pos, (int32_t)b);
    }
    if (pos_extra[b] > 0) { s_put(s, dist - pos_base[b], pos_extra[b]); }
}

static void code_tokens(struct sqz* s, const uint32_t* tokens, uint64_t count) {
    for (uint64_t k = 0; k < count && s->error == 0; k++) {
        const uint32_t t = tokens[k];
        const uint32_t len = t >> 16;
        if (len == 0) {
            code_lit(s, t & 0xFF);
        } else {
            const uint32_t dist = t & 0xFFFF;
            if (len < sqz_min_len || len > sqz_max_len || dist == 0 || dist > 0x7FFF) {
                s->error = EINVAL;      /* the decoder would reject it: squeeze.h:529-545 */
                break;
            }
            code_len(s, len);           /* length first, then distance: squeeze.h:379-380 */
            code_dist(s, dist);
            s->matches++;
        }
    }
    s->tokens += count;
}

void sqz_write_header(struct sqz_bitstream* bs, uint64_t bytes, uint8_t win_bits) {
    if (win_bits < sqz_min_win_bits || win_bits > sqz_max_win_bits) {
        bs->error = EINVAL;
        return;
    }
    put_bits(bs, bytes, 64);
    put_bits(bs, win_bits, 8);
}

void sqz_read_header(struct sqz_bitstream* bs, uint64_t* bytes, uint8_t* win_bits) {
    uint64_t b = get_bits(bs, 64);
    uint64_t w = get_bits(bs, 8);
    if (bs->error != 0) { return; }
    if (w < sqz_min_win_bits || w > sqz_max_win_bits) { bs->error = EINVAL; return; }
    *bytes = b;
    *win_bits = (uint8_t)w;
}

void sqz_encode_tokens(struct sqz* s, struct sqz_bitstream* bs,
                       const uint32_t* tokens, uint64_t count) {
    coder_begin(s, bs);
    double t0 = now_seconds();
    code_tokens(s, tokens, count);
    if (s->error == 0) { pad_to_word(bs); s->error = bs->error; }
    s->entropy_seconds += now_seconds() - t0;
}

void sqz_compress(struct sqz* s, struct sqz_bitstream* bs,
                  const uint8_t* data, uint64_t bytes, uint32_t window) {
    if (window < (1u << sqz_min_win_bits) || window > (1u << sqz_max_win_bits) ||
        (window & (window - 1)) != 0) {
        s->error = EINVAL;
        return;
    }
    coder_begin(s, bs);
    if (s->error != 0) { return; }
    /* The GPU produces the greedy token stream chunk by chunk; while the host
     * entropy-codes chunk k the device already searches chunk k+1.           */
    sqz_gpu_stream* st = NULL;
    double t_search = 0, t_code = 0, t0 = now_seconds();
    int r = sqz_gpu_stream_open(&st, s->device, data, (size_t)bytes, window,
                                sqz_min_len, sqz_max_len, window - 1, 0);
    t_search += now_seconds() - t0;
    if (r != 0) { s->error = r; return; }
    for (;;) {
        const uint32_t* tokens = NULL;
        size_t count = 0;
        t0 = now_seconds();
        r = sqz_gpu_stream_next(st, &tokens, &count);
        t_search += now_seconds() - t0;
        if (r != 0) { s->error = r; break; }
        if (count == 0) { break; }
        t0 = now_seconds();
        code_tokens(s, tokens, count);
        t_code += now_seconds() - t0;
        if (s->error != 0) { break; }
    }
    sqz_gpu_stream_close(st);
    if (s->error == 0) { pad_to_word(bs); s->error = bs->error; }
    s->search_seconds = t_search;
    s->entropy_seconds = t_code;
}

/* ======================================================================== *
 *  decompressor  (reference squeeze.h:411-551); never touches the GPU       *
 * ======================================================================== */

static int32_t read_symbol(struct sqz* s, struct sqz_tree* t) { /* squeeze.h:429-442 */
    const struct sqz_node* nd = t->node;
    int32_t i = tree_root(t);
    for (;;) {
        int bit = get_bit(s->bs);
        if (s->bs->error != 0) { s->error = s->bs->error; return -1; }
        i = bit ? nd[i].hi : nd[i].lo;
        if (i < 0) { s->error = EINVAL; return -1; }
        if (nd[i].lo < 0 && nd[i].hi < 0) { break; }
    }
    tree_count(t, i);
    return i;
}

static inline uint64_t s_get(struct sqz* s, int count) {
    uint64_t v = 0;
    if (s->error == 0) { v = get_bits(s->bs, count); s->error = s->bs->error; }
    return v;
}

void sqz_decompress(struct sqz* s, struct sqz_bitstream* bs,
                    uint8_t* data, uint64_t bytes) {
    coder_begin(s, bs);
    uint64_t i = 0;
    while (i < bytes && s->error == 0) {
        int32_t sym = read_symbol(s, &s->lit);
        if (s->error != 0) { break; }
        if (sym == sqz_lit_nyt) {
            sym = (int32_t)s_get(s, 9);
            if (s->error != 0) { break; }
            if (s->lit.node[sym].up >= 0) { s->error = EINVAL; break; }  /* already known */
            if (!tree_insert(&s->lit, sym)) { s->error = E2BIG; break; }
        }
        if (sym <= 0xFF) {
            data[i++] = (uint8_t)sym;
            continue;
        }
        const int32_t b = sym - len_symbol0;
        if (b < 0 || b >= 28) { s->error = EINVAL; break; }
        uint32_t len = len_base[b];
        if (len_extra[b] > 0) { len += (uint32_t)s_get(s, len_extra[b]); }
        if (s->error != 0) { break; }
        if (len < sqz_min_len || len > sqz_max_len) { s->error = EINVAL; break; }
        int32_t pb = read_symbol(s, &s->pos);
        if (s->error != 0) { break; }
        if (pb == sqz_pos_nyt) {
            pb = (int32_t)s_get(s, 5);
            if (s->error != 0) { break; }
            if (pb >= 30 || s->pos.node[pb].up >= 0) { s->error = EINVAL; break; }
            if (!tree_insert(&s->pos, pb)) { s->error = E2BIG; break; }
        }
        if (pb >= 30) { s->error = EINVAL; break; }
        uint32_t dist = pos_base[pb];
        if (pos_extra[pb] > 0) { dist += (uint32_t)s_get(s, pos_extra[pb]); }
        if (s->error != 0) { break; }
        if (dist == 0 || dist > 0x7FFF || dist > i || len > bytes - i) {
            s->error = EINVAL;
            break;
        }
        /* the source may overlap the destination: copy forward byte by byte */
        const uint8_t* from = data + i - dist;
        for (uint32_t k = 0; k < len; k++) { data[i + k] = from[k]; }
        i += len;
    }
}

/* ---- whole-buffer conveniences ----------------------------------------- */

static struct sqz* state_new(void);
static void state_free(struct sqz* s);
#include <stdlib.h>
static struct sqz* state_new(void) { return (struct sqz*)malloc(sizeof(struct sqz)); }
static void state_free(struct sqz* s) { free(s); }

int sqz_compress_buffer(const uint8_t* data, uint64_t bytes, uint8_t win_bits,
                        uint8_t* out, uint64_t capacity, uint64_t* written) {
    struct sqz_bitstream bs;
    memset(&bs, 0, sizeof(bs));
    bs.data = out;
    bs.capacity = capacity;
    sqz_write_header(&bs, bytes, win_bits);
    if (bs.error != 0) { return bs.error; }
    struct sqz* s = state_new();
    if (s == NULL) { return ENOMEM; }
    sqz_init(s);
    sqz_compress(s, &bs, data, bytes, 1u << win_bits);
    int r = s->error;
    if (written != NULL) { *written = bs.bytes; }
    state_free(s);
    return r;
}

int sqz_decompress_buffer(const uint8_t* comp, uint64_t comp_bytes,
                          uint8_t* out, uint64_t capacity, uint64_t* bytes) {
    struct sqz_bitstream bs;
    memset(&bs, 0, sizeof(bs));
    bs.data = (uint8_t*)comp;
    bs.capacity = comp_bytes;
    bs.bytes = comp_bytes;
    uint64_t n = 0;
    uint8_t wb = 0;
    sqz_read_header(&bs, &n, &wb);
    if (bs.error != 0) { return bs.error; }
    if (n > capacity) { return E2BIG; }
    struct sqz* s = state_new();
    if (s == NULL) { return ENOMEM; }
    sqz_init(s);
    sqz_decompress(s, &bs, out, n);
    int r = s->error;
    if (bytes != NULL) { *bytes = n; }
    state_free(s);
    return r;
}
