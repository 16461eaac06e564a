/* oracle/ref_bst_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Compiles the reference's stand-alone brute-force finder, lz77_find() in
 * /root/reference/bst.c:230-252 (rule set (iii) of SURVEY.md section 8a:
 * min 2, max 254, dist <= window), UNMODIFIED, by including bst.c by path with
 * its main() renamed.  Used to pin the parameterised search (min_len, max_len,
 * max_dist as runtime arguments) against a second, independently written
 * reference loop.  Flags: see oracle/Makefile.
 */
#include <stddef.h>
#include <stdint.h>
#include "bst.c"

__attribute__((visibility("default")))
void ref_bst_lz77_find(const uint8_t* data, uint64_t bytes, uint64_t i,
                       uint64_t window, uint64_t* size, uint64_t* dist) {
    static struct sqz s; /* 2 MiB of tree nodes we never touch */
    s.window = (size_t)window;
    size_t sz = 0, ds = 0;
    lz77_find(&s, data, (size_t)bytes, (size_t)i, &sz, &ds);
    *size = sz; *dist = ds;
}

/* every position at once (positions are independent) */
__attribute__((visibility("default")))
void ref_bst_table(const uint8_t* data, uint64_t bytes, uint64_t window,
                   uint64_t first, uint64_t count,
                   uint16_t* len_out, uint16_t* dist_out) {
    static struct sqz s;
    s.window = (size_t)window;
    for (uint64_t k = 0; k < count; k++) {
        size_t sz = 0, ds = 0;
        lz77_find(&s, data, (size_t)bytes, (size_t)(first + k), &sz, &ds);
        len_out[k] = (uint16_t)sz; dist_out[k] = (uint16_t)ds;
    }
}
