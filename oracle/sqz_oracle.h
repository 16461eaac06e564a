/* oracle/sqz_oracle.h -- TEST INFRASTRUCTURE (the checker), never shipped.
 * Plain-C restatement of the reference's longest-match search and greedy parse
 * (/root/reference/attic/map_experiment/squeeze.h:337-395).  See sqz_oracle.c.
 */
#ifndef SQZ_ORACLE_H
#define SQZ_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The three rule sets that exist in the reference snapshot:
 *   G1 default : min_len 3, max_len 257, max_dist window-1   squeeze.h:13-15,342
 *   HEAD (off) : min_len 2, max_len 254, max_dist window-1   src/sqz.c:29-30,637-654
 *   bst.c      : min_len 2, max_len 254, max_dist window     bst.c:3,230-252   */
typedef struct {
    uint32_t min_len;
    uint32_t max_len;
    uint32_t max_dist;
} oracle_rules;

/* best (len, dist) at one position; (0,0) when nothing reaches min_len */
void oracle_best(const uint8_t* data, uint64_t bytes, uint64_t i,
                 const oracle_rules* r, uint32_t* len, uint32_t* dist);

/* the loop above evaluated for positions [first, first+count) (OpenMP) */
void oracle_match_table(const uint8_t* data, uint64_t bytes,
                        const oracle_rules* r, uint64_t first, uint64_t count,
                        uint16_t* len_out, uint16_t* dist_out);

/* greedy parse, searching only at parse positions like the reference does;
 * token = literal byte (bits 31..16 zero) or (len << 16) | dist.
 * returns the token count (may exceed cap; only cap are stored)             */
uint64_t oracle_tokens(const uint8_t* data, uint64_t bytes,
                       const oracle_rules* r, uint32_t* tokens, uint64_t cap);

/* greedy walk over an already computed table */
uint64_t oracle_tokens_from_table(const uint8_t* data, uint64_t bytes,
                                  const uint16_t* len, const uint16_t* dist,
                                  uint32_t min_len, uint64_t start,
                                  uint32_t* tokens, uint64_t cap,
                                  uint64_t* end_pos);

/* "oracle B": exact hash-chain search (same results, ~1000x faster) for sizes
 * the brute force cannot finish; validated against oracle_match_table in tests */
void oracle_fast_table(const uint8_t* data, uint64_t bytes,
                       const oracle_rules* r, uint64_t first, uint64_t count,
                       uint16_t* len_out, uint16_t* dist_out);

/* worker threads for the table functions (0 = all online cores) */
void oracle_set_threads(int n);
int  oracle_get_threads(void);

/* FNV-1a 64 over a byte range: the digest the golden fixtures are stored as */
uint64_t oracle_fnv1a64(const void* p, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif
