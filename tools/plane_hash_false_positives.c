// False candidates of a three-byte match test on hashed bytes (6 or 7 of 8 bit planes, match_bitsliced.cuh: kQuietPlanes) per position,
// window 32767:   gcc -O2 -o fp tools/plane_hash_false_positives.c && ./fp FILE 0x2b 0x16 6
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
static uint32_t *ch, *cr;
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* d = malloc(n); fread(d, 1, n, f); fclose(f);
    int a = strtol(argv[2], 0, 0), b = strtol(argv[3], 0, 0);
    int planes = argc > 4 ? atoi(argv[4]) : 6;
    uint8_t T[4] = {0, a, b, a ^ b};
    uint8_t* h = malloc(n);
    for (long i = 0; i < n; i++) {
        if (planes == 6) h[i] = (d[i] & 0x3F) ^ T[d[i] >> 6];
        else if (planes == 7) h[i] = (d[i] & 0x7F) ^ ((d[i] >> 7) ? a : 0);
        else h[i] = d[i];
    }
    ch = calloc(1 << 24, 4); cr = calloc(1 << 24, 4);
    const long W = 32767;
    double fp = 0, tp = 0;
    for (long i = 0; i + 2 < n; i++) {
        uint32_t kh = h[i] | h[i+1] << 8 | h[i+2] << 16, kr = d[i] | d[i+1] << 8 | d[i+2] << 16;
        fp += ch[kh] - cr[kr]; tp += cr[kr];
        ch[kh]++; cr[kr]++;
        if (i >= W) { long j = i - W; ch[h[j] | h[j+1] << 8 | h[j+2] << 16]--; cr[d[j] | d[j+1] << 8 | d[j+2] << 16]--; }
    }
    printf("%s a=0x%02x b=0x%02x planes=%d: true 3-byte candidates/pos %.3f, false %.4f\n", argv[1], a, b, planes, tp / n, fp / n);
    return 0;
}
