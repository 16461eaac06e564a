/* Fuzz of the decoder on damaged streams (not part of the suite, no GPU): a token stream (symbol words
 * as a file of uint32, e.g. np.asarray(sq.symbols_of_tokens(tokens), np.uint32).tofile(...)) is encoded,
 * then decoded `iterations` times with flipped bits, overwritten stretches and cuts -- built with the
 * sanitizers, what counts is that none of them reports anything:
 *   gcc -std=gnu11 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -w -Iinclude \
 *       tools/fuzz_decoder.c -o /tmp/fuzz_decoder -lpthread
 *   ASAN_OPTIONS=detect_leaks=0 /tmp/fuzz_decoder words.bin [tokens] [iterations]                    */
#define SQZ_EXPORT
#include "../sqz_b200/csrc/sqz_codec.c"
#include <stdio.h>
int  sqz_gpu_stream_open(sqz_gpu_stream** st, int device, const uint8_t* d, size_t n, uint32_t w, uint32_t a, uint32_t b, uint32_t c, size_t e, uint32_t f) { return ENODEV; }
int  sqz_gpu_stream_next(sqz_gpu_stream* st, const uint32_t** tokens, size_t* count) { return ENODEV; }
void sqz_gpu_stream_close(sqz_gpu_stream* st) {}
int sqz_gpu_expand_tokens(const uint32_t* tokens, size_t n_tokens, uint8_t* out, size_t bytes) { return ENODEV; }
static uint64_t rs = 88172645463325252ull;
static uint64_t rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; }
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); size_t count = ftell(f) / 4; fseek(f, 0, SEEK_SET);
    if (argc > 2 && (size_t)atol(argv[2]) < count) count = atol(argv[2]);
    int iters = argc > 3 ? atoi(argv[3]) : 500;
    uint32_t* words = malloc(count * 4); if (fread(words, 4, count, f) != count) return 1; fclose(f);
    size_t cap = count * 9 + 64; uint8_t* out = malloc(cap); uint8_t* bad = malloc(cap);
    struct sqz* s = malloc(sizeof(struct sqz));
    struct sqz_bitstream bs; memset(&bs, 0, sizeof bs); bs.data = out; bs.capacity = cap;
    sqz_init(s); s->coder_threads = 1; sqz_encode_symbols(s, &bs, words, count);
    uint64_t nbytes = 0; for (size_t k = 0; k < count; k++) { uint32_t sym = words[k] & 0x1FF; if (sym < 256) nbytes++; else { uint32_t b = sym - 257; uint32_t lx = (words[k] >> 9) & 31; nbytes += len_base[b] + reverse_field(lx, len_extra[b]); } }
    uint8_t* data = malloc(nbytes + 1);
    int errors = 0, clean = 0;
    for (int it = 0; it < iters; it++) {
        memcpy(bad, out, bs.bytes);
        size_t len = bs.bytes;
        int kind = rnd() % 4;
        if (kind == 0) { for (int q = 0; q < 1 + (int)(rnd() % 3); q++) bad[rnd() % len] ^= (uint8_t)(1u << (rnd() % 8)); }
        else if (kind == 1) { len = (rnd() % len) & ~(size_t)7; }                       /* cut at a word */
        else if (kind == 2) { size_t a = rnd() % len; size_t n = 1 + rnd() % 64; for (size_t q = a; q < a + n && q < len; q++) bad[q] = (uint8_t)rnd(); }
        else { len = rnd() % len; }                                                      /* cut anywhere */
        struct sqz_bitstream rd; memset(&rd, 0, sizeof rd); rd.data = bad; rd.bytes = len; rd.capacity = cap;
        sqz_init(s);
        if (it & 1) { sqz_decompress(s, &rd, data, nbytes); } else { uint64_t got = 0; sqz_decode_tokens(s, &rd, nbytes, NULL, 0, &got); }
        if (s->error) errors++; else clean++;
    }
    printf("%d damaged streams: %d reported an error, %d decoded to something\n", iters, errors, clean);
    return 0;
}
