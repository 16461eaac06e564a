/* tests/c/gpu_consumer.c -- TEST INFRASTRUCTURE: a C program that consumes libsqz_b200.so the way
 * INTEGRATION.md section 1 tells a maintainer of leok7v/sqz to.
 *
 * It includes the UNMODIFIED reference codec by path
 * (/root/reference/attic/map_experiment/{bitstream,huffman,squeeze}.h; nothing is copied) and
 * defines gpu_squeeze_compress(): squeeze_compress (squeeze.h:319-409) with its search loop
 * (squeeze.h:338-358) and greedy dispatch (squeeze.h:377-394) replaced by
 * sqz_gpu_stream_open / sqz_gpu_stream_next / sqz_gpu_stream_close and a token loop around the
 * reference's own squeeze_encode_literal / squeeze_encode_len / squeeze_encode_pos.  Everything
 * else -- the adaptive Huffman trees, the bitstream, the header -- is the reference's own code.
 * For every file named on the command line it compresses twice, once with the reference's
 * squeeze.compress (CPU search) and once with gpu_squeeze_compress, and compares the bytes.
 *
 * Built by oracle/Makefile (target ref) into oracle/_ref/gpu_consumer, linked with -lsqz_b200;
 * run by tests/test_c_consumer.py on the GPU box.
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

#include "bitstream.h"
#define squeeze_implementation
#include "squeeze.h"

#include "sqz_gpu.h"

static void gpu_squeeze_compress(squeeze_type* s, bitstream* bs, const uint8_t* data, uint64_t bytes,
                                 uint16_t window) {
    s->bs = bs;
    if (!huffman_insert(&s->lit, squeeze_lit_nyt)) { s->error = EINVAL; }
    if (!huffman_insert(&s->pos, squeeze_pos_nyt)) { s->error = EINVAL; }
    squeeze_deflate_init(s);
    sqz_gpu_stream* st = NULL;
    int r = sqz_gpu_stream_open(&st, 0, data, bytes, window, squeeze_deflate_len_min, squeeze_deflate_len_max,
                                window - 1u, 0, 0);
    if (r != 0) { s->error = r; return; }
    for (;;) {
        const uint32_t* tok = NULL;
        size_t n = 0;
        r = sqz_gpu_stream_next(st, &tok, &n);
        if (r != 0) { s->error = r; break; }
        if (n == 0) { break; }
        for (size_t k = 0; k < n && s->error == 0; k++) {
            if ((tok[k] >> 16) == 0) {
                squeeze_encode_literal(s, (uint16_t)tok[k]);
            } else {
                squeeze_encode_len(s, (uint16_t)(tok[k] >> 16));
                squeeze_encode_pos(s, (uint16_t)(tok[k] & 0xFFFF));
            }
        }
    }
    sqz_gpu_stream_close(st);
    squeeze_flush(s);
}

static int compress_with(int gpu, const uint8_t* data, uint64_t bytes, uint8_t* out, uint64_t cap, uint64_t* written) {
    bitstream bs;
    memset(&bs, 0, sizeof(bs));
    bs.data = out;
    bs.capacity = cap;
    squeeze.write_header(&bs, bytes, 15);
    if (bs.error != 0) { return bs.error; }
    squeeze_type* s = squeeze.alloc(0);
    if (s == null) { return ENOMEM; }
    if (gpu) { gpu_squeeze_compress(s, &bs, data, bytes, (uint16_t)(1u << 15)); }
    else     { squeeze.compress(s, &bs, data, bytes, (uint16_t)(1u << 15)); }
    int r = s->error;
    *written = bs.bytes;
    squeeze.free(s);
    return r;
}

int main(int argc, char** argv) {
    if (sqz_gpu_device_count() < 1) { fprintf(stderr, "no CUDA device\n"); return 2; }
    int bad = 0;
    for (int a = 1; a < argc; a++) {
        FILE* f = fopen(argv[a], "rb");
        if (f == NULL) { perror(argv[a]); return 2; }
        fseek(f, 0, SEEK_END);
        long size = ftell(f);
        fseek(f, 0, SEEK_SET);
        uint8_t* data = (uint8_t*)malloc((size_t)size + 1);
        if (fread(data, 1, (size_t)size, f) != (size_t)size) { perror("fread"); return 2; }
        fclose(f);
        uint64_t cap = (uint64_t)size * 2 + 4096, wa = 0, wb = 0;
        uint8_t* A = (uint8_t*)calloc(cap, 1);
        uint8_t* B = (uint8_t*)calloc(cap, 1);
        int ra = compress_with(0, data, (uint64_t)size, A, cap, &wa);
        int rb = compress_with(1, data, (uint64_t)size, B, cap, &wb);
        int same = ra == 0 && rb == 0 && wa == wb && memcmp(A, B, wa) == 0;
        printf("%s: %ld bytes -> reference %llu (rc %d), GPU search + reference encoder %llu (rc %d%s%s): %s\n",
               argv[a], size, (unsigned long long)wa, ra, (unsigned long long)wb, rb,
               rb != 0 ? ", " : "", rb != 0 ? sqz_gpu_last_error() : "", same ? "identical" : "DIFFERENT");
        bad += !same;
        free(data); free(A); free(B);
    }
    return bad ? 1 : 0;
}
