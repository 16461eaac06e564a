/* oracle/ref_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Compiles the UNMODIFIED reference codec (generation G1,
 * /root/reference/attic/map_experiment/{bitstream,huffman,map,squeeze}.h) into
 * oracle/_ref/libsqzref.so by including the headers BY PATH (-I); no reference
 * source is copied into this repository.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load the result.
 *
 * What it exports ("oracle A" of SURVEY.md section 8c):
 *   ref_compress        squeeze.write_header + squeeze.compress, memory mode
 *   ref_compress_cb     same through the 8-byte output callback (file mode)
 *   ref_decompress      squeeze.read_header + squeeze.decompress
 *   ref_tokens          the reference's own (len,pos)/literal decisions, read
 *                       back from its bitstream with its own static decoder
 *                       functions: a host-side dump of the reference search at
 *                       every parse position
 *   ref_encode_tokens   a caller-supplied token list pushed through the
 *                       reference's own squeeze_encode_{literal,len,pos}
 *
 * The prelude below supplies what the reference headers use without
 * including: errno_t, null, bool, memset, log2.
 */
typedef int errno_t;
#define null ((void*)0)
#include <assert.h>
#include <errno.h>
#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "bitstream.h"
#define squeeze_implementation
#include "squeeze.h"

#define REF_API __attribute__((visibility("default")))

static double ref_now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* token word shared with the product: literal = byte value (bits 31..16 zero),
 * match = (len << 16) | dist                                                 */
static inline uint32_t ref_tok_match(uint32_t len, uint32_t dist) {
    return (len << 16) | dist;
}

REF_API int ref_compress(const uint8_t* data, uint64_t bytes, int win_bits,
                         uint8_t* out, uint64_t cap, uint64_t* written,
                         double* seconds) {
    bitstream bs;
    memset(&bs, 0, sizeof(bs));
    bs.data = out;
    bs.capacity = cap;
    squeeze.write_header(&bs, bytes, (uint8_t)win_bits);
    if (bs.error != 0) { return bs.error; }
    squeeze_type* s = squeeze.alloc(0);
    if (s == null) { return ENOMEM; }
    double t0 = ref_now();
    squeeze.compress(s, &bs, data, bytes, (uint16_t)(1u << win_bits));
    double t1 = ref_now();
    int r = s->error;
    if (written != null) { *written = bs.bytes; }
    if (seconds != null) { *seconds = t1 - t0; }
    squeeze.free(s);
    return r;
}

typedef struct { uint8_t* out; uint64_t cap; uint64_t at; } ref_sink;

static errno_t ref_sink_output(bitstream* bs) {
    ref_sink* k = (ref_sink*)bs->stream;
    if (k->at + 8 > k->cap) { return E2BIG; }
    memcpy(k->out + k->at, &bs->b64, 8); /* what fwrite(&bs->b64, 8, 1, f) stores */
    k->at += 8;
    return 0;
}

REF_API int ref_compress_cb(const uint8_t* data, uint64_t bytes, int win_bits,
                            uint8_t* out, uint64_t cap, uint64_t* written) {
    ref_sink sink = { out, cap, 0 };
    bitstream bs;
    memset(&bs, 0, sizeof(bs));
    bs.stream = &sink;
    bs.output = ref_sink_output;
    squeeze.write_header(&bs, bytes, (uint8_t)win_bits);
    if (bs.error != 0) { return bs.error; }
    squeeze_type* s = squeeze.alloc(0);
    if (s == null) { return ENOMEM; }
    squeeze.compress(s, &bs, data, bytes, (uint16_t)(1u << win_bits));
    int r = s->error;
    if (written != null) { *written = bs.bytes; }
    squeeze.free(s);
    return r;
}

REF_API int ref_read_header(const uint8_t* comp, uint64_t comp_bytes,
                            uint64_t* bytes, int* win_bits) {
    bitstream bs;
    memset(&bs, 0, sizeof(bs));
    bs.data = (uint8_t*)comp;
    bs.capacity = comp_bytes;
    bs.bytes = comp_bytes;
    uint64_t b = 0; uint8_t w = 0;
    squeeze.read_header(&bs, &b, &w);
    if (bs.error == 0) { *bytes = b; *win_bits = w; }
    return bs.error;
}

REF_API int ref_decompress(const uint8_t* comp, uint64_t comp_bytes,
                           uint8_t* out, uint64_t cap, uint64_t* bytes,
                           double* seconds) {
    bitstream bs;
    memset(&bs, 0, sizeof(bs));
    bs.data = (uint8_t*)comp;
    bs.capacity = comp_bytes;
    bs.bytes = comp_bytes;
    uint64_t b = 0; uint8_t w = 0;
    squeeze.read_header(&bs, &b, &w);
    if (bs.error != 0) { return bs.error; }
    if (b > cap) { return E2BIG; }
    squeeze_type* s = squeeze.alloc(0);
    if (s == null) { return ENOMEM; }
    double t0 = ref_now();
    squeeze.decompress(s, &bs, out, b);
    double t1 = ref_now();
    int r = s->error;
    if (bytes != null) { *bytes = b; }
    if (seconds != null) { *seconds = t1 - t0; }
    squeeze.free(s);
    return r;
}

/* Walk the reference's bitstream with the reference's own static readers and
 * report the decision it took at every parse position.                      */
REF_API int ref_tokens(const uint8_t* comp, uint64_t comp_bytes,
                       uint32_t* tokens, uint64_t cap, uint64_t* count) {
    bitstream bs;
    memset(&bs, 0, sizeof(bs));
    bs.data = (uint8_t*)comp;
    bs.capacity = comp_bytes;
    bs.bytes = comp_bytes;
    uint64_t bytes = 0; uint8_t w = 0;
    squeeze.read_header(&bs, &bytes, &w);
    if (bs.error != 0) { return bs.error; }
    squeeze_type* s = squeeze.alloc(0);
    if (s == null) { return ENOMEM; }
    s->bs = &bs;
    if (!huffman_insert(&s->lit, squeeze_lit_nyt)) { s->error = EINVAL; }
    if (!huffman_insert(&s->pos, squeeze_pos_nyt)) { s->error = EINVAL; }
    squeeze_deflate_init(s);
    uint64_t i = 0, n = 0;
    while (i < bytes && s->error == 0) {
        uint64_t lit = squeeze_read_huffman(s, &s->lit);
        if (s->error != 0) { break; }
        if (lit == squeeze_lit_nyt) {
            lit = squeeze_read_bits(s, 9);
            if (s->error != 0) { break; }
            if (!huffman_insert(&s->lit, (int32_t)lit)) { s->error = E2BIG; break; }
        }
        uint32_t t;
        if (lit <= 0xFF) {
            t = (uint32_t)lit;
            i++;
        } else {
            uint32_t len = squeeze_read_length(s, (uint16_t)lit);
            if (s->error != 0) { break; }
            uint32_t pos = squeeze_read_pos(s);
            if (s->error != 0) { break; }
            t = ref_tok_match(len, pos);
            i += len;
        }
        if (n < cap) { tokens[n] = t; }
        n++;
    }
    int r = s->error;
    if (r == 0 && n > cap) { r = E2BIG; }
    *count = n;
    squeeze.free(s);
    return r;
}

/* Push a token list through the reference's own symbol coder + bit writer.  */
REF_API int ref_encode_tokens(const uint32_t* tokens, uint64_t count,
                              uint64_t bytes, int win_bits, int file_mode,
                              uint8_t* out, uint64_t cap, uint64_t* written,
                              double* seconds) {
    ref_sink sink = { out, cap, 0 };
    bitstream bs;
    memset(&bs, 0, sizeof(bs));
    if (file_mode) { bs.stream = &sink; bs.output = ref_sink_output; }
    else           { bs.data = out; bs.capacity = cap; }
    squeeze.write_header(&bs, bytes, (uint8_t)win_bits);
    if (bs.error != 0) { return bs.error; }
    squeeze_type* s = squeeze.alloc(0);
    if (s == null) { return ENOMEM; }
    double t0 = ref_now();
    s->bs = &bs;
    if (!huffman_insert(&s->lit, squeeze_lit_nyt)) { s->error = EINVAL; }
    if (!huffman_insert(&s->pos, squeeze_pos_nyt)) { s->error = EINVAL; }
    squeeze_deflate_init(s);
    for (uint64_t k = 0; k < count && s->error == 0; k++) {
        uint32_t t = tokens[k];
        if ((t >> 16) == 0) {
            squeeze_encode_literal(s, (uint16_t)(t & 0xFF));
        } else {
            squeeze_encode_len(s, (uint16_t)(t >> 16));
            squeeze_encode_pos(s, (uint16_t)(t & 0xFFFF));
        }
    }
    squeeze_flush(s);
    double t1 = ref_now();
    int r = s->error;
    if (written != null) { *written = bs.bytes; }
    if (seconds != null) { *seconds = t1 - t0; }
    squeeze.free(s);
    return r;
}

REF_API int ref_abi_version(void) { return 1; }
