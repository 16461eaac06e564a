/* include/sqz_gpu.h -- C-ABI of the B200 (sm_100a) LZ77 match search of sqz-b200.
 *
 * The reference (leok7v/sqz) has no FFI or plugin interface: its longest-match
 * search is an inline loop inside squeeze_compress
 * (/root/reference/attic/map_experiment/squeeze.h:338-358) followed by the
 * greedy token dispatch (squeeze.h:377-394).  These entry points are what a
 * maintainer would call at exactly that place: plain
 * pointers and sizes, int errno return values, callable from C99.  No call
 * changes the calling thread's current CUDA device.  State that outlives a call
 * is limited to recycled staging buffers (sqz_gpu_release frees them) and the
 * measurement aids at the end of this file, which are off unless switched on.
 *
 * Rule parameters:
 *   window    the LZ window the stream header announces (power of two)
 *   min_len   shortest match worth a back-reference     (reference G1: 3)
 *   max_len   longest match                              (reference G1: 257)
 *   max_dist  farthest candidate                         (reference G1: window-1)
 * For every position i the result is the candidate with the longest common
 * prefix (capped by max_len and by bytes-i), the nearest one among equals --
 * what the reference's "nearest first, strictly longer wins" scan selects.
 *
 * Token word (shared with sqz.h): literal = byte value (bits 31..16 zero);
 * match = (len << 16) | dist.  A match-table word uses the same packing, with
 * 0 meaning "no match of at least min_len here".
 *
 * Symbol word (what the host coder of sqz.h consumes directly, sqz_encode_symbols): the same
 * token with the bucket arithmetic of squeeze_encode_len / squeeze_encode_pos
 * (squeeze.h:290-315, tables squeeze.h:29-79,151-172) already done on the GPU:
 *   bits  0..8   symbol of the literal/length tree: byte value, or 257 + length bucket
 *   bits  9..13  the length's extra bits      (len - len_base[bucket])
 *   bits 14..18  symbol of the distance tree  (distance bucket)
 *   bits 19..31  the distance's extra bits    (dist - pos_base[bucket])
 * Extra bits are stored in emission order, i.e. bit-reversed within their field
 * width, because the bitstream takes values least significant bit first
 * (bitstream.h:49-63).  Symbol words exist only within the bitstream's own
 * limits: min_len >= 3, max_len <= 257 (what the decoders accept, squeeze.h:529-545),
 * max_dist <= 32767.
 *
 * Errors: 0, EINVAL (bad rule parameters), E2BIG (output capacity), ENOMEM,
 * ENODEV (no CUDA device / driver: there is NO CPU fallback), EIO (CUDA error;
 * sqz_gpu_last_error() has the text).
 */
#ifndef SQZ_GPU_H_INCLUDED
#define SQZ_GPU_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQZ_GPU_ABI_VERSION 3

enum {
    sqz_gpu_max_len_limit  = 512,     /* parse hand-off tables are sized for this */
    sqz_gpu_max_dist_limit = 65535    /* distances are reported in 16 bits */
};

/* ---- host-buffer entry points (replace squeeze.h:338-358 / 377-394) ------ *
 * They run on the calling thread's current CUDA device (device 0 unless the
 * caller selected another one); sqz_gpu_stream_open takes the device explicitly
 * (-1 = current).                                                            */

/* Full match table: for every i in [0, bytes) len_out[i] in {0} U
 * [min_len, max_len] and dist_out[i] in [1, max_dist] (0 when len is 0).
 * Replaces the search loop squeeze.h:340-358 evaluated at every position.   */
int sqz_gpu_match_table(const uint8_t* data, size_t bytes,
                        uint32_t window, uint32_t min_len, uint32_t max_len,
                        uint32_t max_dist,
                        uint16_t* len_out, uint16_t* dist_out);

/* Greedy token stream in parse order (squeeze.h:337,377-394): search + parse
 * on the GPU, tokens copied to the caller.  *n_tokens receives the count even
 * when it exceeds tokens_cap (then E2BIG is returned).                       */
int sqz_gpu_tokens(const uint8_t* data, size_t bytes,
                   uint32_t window, uint32_t min_len, uint32_t max_len,
                   uint32_t max_dist,
                   uint32_t* tokens_out, size_t tokens_cap, size_t* n_tokens);

/* The same token stream computed on several devices at once (SURVEY.md section 8e;
 * squeeze.h:345,377-394 is where the seam dependency comes from).  The input is cut
 * into n_devices contiguous shards (sizes differ by at most one byte), each uploaded
 * with a look-back halo of max_dist bytes and a look-ahead halo of max_len bytes, so
 * the match tables are independent.  The parse depends on one number per seam -- where
 * the previous shard's last token ends -- which every device prepares for as a 1 KiB
 * exit map and the host chains with one lookup per seam.  The token arrays are then
 * concatenated in shard order by copies sized by their counts: device -> host when
 * tokens_out is host memory (pinned memory takes them by DMA from all devices at
 * once), peer-to-peer over NVLink when tokens_out is device memory.  No collective
 * library is involved.  The result is identical to sqz_gpu_tokens() for every
 * n_devices.  A device may be listed more than once (its shards then share it).
 * shard_tokens (optional, n_devices entries) receives the per-shard counts.       */
int sqz_gpu_tokens_multi(const int* devices, int n_devices,
                         const uint8_t* data, size_t bytes,
                         uint32_t window, uint32_t min_len, uint32_t max_len,
                         uint32_t max_dist,
                         uint32_t* tokens_out, size_t tokens_cap, size_t* n_tokens,
                         size_t* shard_tokens);

/* Streaming form used by sqz_compress(): tokens arrive chunk by chunk, in
 * parse order, from double-buffered pinned memory while the device already
 * works on the next chunk.  *tokens stays valid until the next call.
 * chunk_bytes = 0: chunks of 2, 4, 8, 16, then 32 MiB -- a consumer slower
 * than the search starts after milliseconds.  chunk_bytes > 0: chunks of that
 * size -- for a consumer faster than the search, which only has to keep the
 * device on chunks it is efficient at; with SQZ_GPU_STREAM_SHORT_START the
 * first two are a quarter and a half of it.  The searches of consecutive
 * chunks run one after the other (a consumer wants them in order, not both
 * late), a chunk's parse and copy at the highest stream priority beside the
 * next chunk's search.                                                      */
typedef struct sqz_gpu_stream sqz_gpu_stream;
#define SQZ_GPU_STREAM_SYMBOLS 1u       /* deliver symbol words instead of plain tokens */
#define SQZ_GPU_STREAM_SHORT_START 2u   /* explicit chunk_bytes: start with chunk_bytes / 4, then / 2 */
int  sqz_gpu_stream_open(sqz_gpu_stream** st, int device,
                         const uint8_t* data, size_t bytes,
                         uint32_t window, uint32_t min_len, uint32_t max_len,
                         uint32_t max_dist, size_t chunk_bytes /* 0 = default */,
                         uint32_t flags /* SQZ_GPU_STREAM_* */);
int  sqz_gpu_stream_next(sqz_gpu_stream* st, const uint32_t** tokens, size_t* count);
void sqz_gpu_stream_close(sqz_gpu_stream* st);

/* ---- device-buffer entry points (shards, benchmarks, multi-GPU) ---------- *
 * A shard is `n` positions starting at d_shard, with `back` valid bytes
 * before it (look-back halo, min(global offset, max_dist) is enough) and
 * `ahead` valid bytes after it (look-ahead halo, min(bytes to the global end,
 * max_len) is enough).  All pointers are device pointers on the current
 * device; `stream` is a cudaStream_t (NULL = default stream).  Asynchronous. */

/* table[i] for the n positions of the shard, packed (len << 16) | dist.
 * The kernels need sqz_gpu_match_workspace(n) bytes of scratch device memory (the
 * work list the first kernel leaves for the second): the _ws form takes it from the
 * caller, the plain form borrows it from a per-device pool of the library.   */
size_t sqz_gpu_match_workspace(size_t n);
int sqz_gpu_match_table_device_ws(const uint8_t* d_shard, size_t back, size_t n,
                                  size_t ahead, uint32_t min_len, uint32_t max_len,
                                  uint32_t max_dist, uint32_t* d_table, void* d_work,
                                  void* stream);
int sqz_gpu_match_table_device(const uint8_t* d_shard, size_t back, size_t n,
                               size_t ahead, uint32_t min_len, uint32_t max_len,
                               uint32_t max_dist, uint32_t* d_table, void* stream);

/* Split a packed table into the two 16-bit arrays of sqz_gpu_match_table.   */
int sqz_gpu_unpack_table_device(const uint32_t* d_table, size_t n,
                                uint16_t* d_len, uint16_t* d_dist, void* stream);

/* Greedy parse of one shard.  `entry` is the offset of the first parse
 * position inside the shard (0 for the first shard; otherwise the overshoot
 * of the previous shard's last token).  d_work must hold
 * sqz_gpu_parse_workspace(n) bytes.  After the stream is synchronised,
 * h_result[0] = token count, h_result[1] = overshoot into the next shard
 * (h_result: 2 x uint64 in pinned or pageable host memory, or device memory
 * when result_on_device != 0).  Tokens beyond tokens_cap are not stored.     */
size_t sqz_gpu_parse_workspace(size_t n);
int sqz_gpu_parse_device(const uint8_t* d_shard, const uint32_t* d_table, size_t n,
                         uint32_t entry, uint32_t min_len, uint32_t max_len,
                         uint32_t* d_tokens, size_t tokens_cap,
                         void* d_work, uint64_t* d_result /* 2 x u64, device */,
                         void* stream);

/* The same parse, emitting symbol words (see above) instead of plain tokens:
 * replaces squeeze.h:377-394 together with the table lookups and subtractions
 * of squeeze.h:290-315.                                                      */
int sqz_gpu_parse_symbols_device(const uint8_t* d_shard, const uint32_t* d_table, size_t n,
                                 uint32_t entry, uint32_t min_len, uint32_t max_len,
                                 uint32_t* d_words, size_t words_cap,
                                 void* d_work, uint64_t* d_result /* 2 x u64, device */,
                                 void* stream);

/* Seam hand-off without waiting for the previous shard: exit_map[e] is the
 * overshoot this shard produces when entered at offset e, for every
 * e < max_len.  d_exit_map: max_len x uint16 on the device.                  */
int sqz_gpu_parse_exit_map_device(const uint32_t* d_table, size_t n,
                                  uint32_t min_len, uint32_t max_len,
                                  void* d_work, uint16_t* d_exit_map, void* stream);

/* ---- one process per device ------------------------------------------------ *
 * Jobs that run one process per GPU (torchrun and the like) concatenate their
 * shards' tokens the same way as sqz_gpu_tokens_multi, except that the destination
 * buffer lives in one process and has to be mapped into the others: CUDA IPC, no
 * collective.  The owner allocates and exports, the others open; every process then
 * puts its tokens at its offset (the sum of the earlier shards' counts).       */
#define SQZ_GPU_IPC_HANDLE_BYTES 64
int  sqz_gpu_device_alloc(void** d_ptr, size_t bytes);
void sqz_gpu_device_free(void* d_ptr);
int  sqz_gpu_ipc_export(const void* d_ptr, uint8_t handle[SQZ_GPU_IPC_HANDLE_BYTES]);
int  sqz_gpu_ipc_open(const uint8_t handle[SQZ_GPU_IPC_HANDLE_BYTES], void** d_ptr);
int  sqz_gpu_ipc_close(void* d_ptr);
int  sqz_gpu_put_tokens(uint32_t* d_dst, size_t at, const uint32_t* d_tokens, size_t count, void* stream);

/* ---- decoder side: the LZ copy phase (replaces squeeze.h:533-539) --------- *
 * The reference's decoder executes every token as it reads it: a literal is
 * stored, a match copies len bytes from dist back, one byte at a time because
 * the ranges may overlap.  With the tokens at hand (sqz_decode_tokens in sqz.h
 * reads them from the bitstream on the host, which is the serial part) the
 * copies are a parallel problem: a scan of the token lengths places every
 * token, every output byte starts with a hop of `dist` to its source, and
 * pointer doubling shortens all chains to one hop onto a literal.  Plain
 * tokens (not symbol words); bytes < 2 GiB per call.  EINVAL when the tokens
 * do not describe exactly `bytes` bytes or a match reaches before the start. */
int sqz_gpu_expand_tokens(const uint32_t* tokens, size_t n_tokens, uint8_t* out, size_t bytes);
/* device pointers; d_work holds sqz_gpu_expand_workspace(n_tokens, bytes) bytes.
 * Waits for the stream twice: once for the scan of the token lengths (a stream
 * that does not describe `bytes` bytes is refused before anything is placed),
 * once at the end for the verdict on the distances.  The rounds of pointer
 * doubling in between end themselves on the device.                           */
size_t sqz_gpu_expand_workspace(size_t n_tokens, size_t bytes);
int sqz_gpu_expand_tokens_device(const uint32_t* d_tokens, size_t n_tokens, uint8_t* d_out,
                                 size_t bytes, void* d_work, void* stream);

/* ---- utilities ----------------------------------------------------------- */
int         sqz_gpu_abi_version(void);
int         sqz_gpu_device_count(void);          /* 0 when no driver / device */
const char* sqz_gpu_last_error(void);            /* thread-local text */
void*       sqz_gpu_host_alloc(size_t bytes);    /* pinned host memory (NULL on failure) */
void        sqz_gpu_host_free(void* p);
/* The host-buffer entry points keep their device and pinned staging buffers
 * for the next call (at most four slots and 3 GiB, oldest out first), and the
 * plain device entry point keeps its scratch buffers; this frees them all.   */
void        sqz_gpu_release(void);
/* Which match-table kernel serves sqz_gpu_match_table_device: 0 = automatic
 * (bit-sliced kernel for min_len 2 or 3, thread-per-position kernel otherwise),
 * 1 = thread-per-position, 2 = bit-sliced where applicable.  Both are exact;
 * the switch exists for A/B measurements and tests and applies to the calling
 * thread only.                                                               */
int         sqz_gpu_select_kernel(int which);
/* Debugging aid: when d_buf (device, one u64 per tile of 16256 positions) is not
 * NULL the bit-sliced kernel stores every tile's duration in SM cycles.      */
void        sqz_gpu_debug_tile_cycles(unsigned long long* d_buf);
/* kernels launched by this library in this process so far (for bench.py)    */
uint64_t    sqz_gpu_launch_count(void);
/* average device time of the match-table kernel over the launches since the
 * last call with reset != 0, measured with CUDA events on the launching
 * stream; enable with sqz_gpu_set_timing(1).  Seconds.                       */
void        sqz_gpu_set_timing(int on);
double      sqz_gpu_match_kernel_seconds(int reset, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* SQZ_GPU_H_INCLUDED */
