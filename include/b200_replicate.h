/* b200_replicate.h — C ABI of libb200mtm.so: replicating an operand across the GPUs of one
 * NVSwitch box with this library's own kernels (multi-GPU row-block mtm).
 *
 * The reference is single-process shared memory: every OpenMP thread of its M-block loop reads the
 * same packed B panel (`buffB`, include/mtm.hpp:151,168-176).  Across GPUs the analogue is a copy
 * of B in every GPU's HBM; these entry points move it there:
 *
 *   root:      b200_flag_wait (every receiver has finished with the buffer)  ->
 *              per K-chunk: b200_replicate_push (data, then the chunk's arrival flag)
 *   receiver:  per K-chunk: b200_flag_wait (arrival flag) -> b200_mtm_*_dev on the chunk;
 *              after the last chunk: b200_flag_signal (to the root's flag word of this rank)
 *
 * All addresses are device-accessible virtual addresses the caller maps (peer-mapped symmetric
 * allocations; a multicast address covers the same offset of every GPU's buffer).  Flags are
 * uint32 sequence numbers compared wrap-safe (a flag "has reached" v when (int32)(flag - v) >= 0).
 * Pushes may overlap on different streams (each launch has its own completion counter).  Status codes as in
 * b200_mtm.h.
 */
#ifndef B200_REPLICATE_H
#define B200_REPLICATE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Copy `bytes` (multiple of 16; src and every dst 16-byte aligned) from `src` to dst[0..n_dst).
 * multicast != 0: n_dst must be 1 and dst[0] a multicast address (multimem.st: the switch
 * replicates each store); otherwise dst[i] are ordinary (peer) addresses written in turn.
 * After all data stores are fenced system-wide, `flag_value` is written to flag_dst[0..n_flag_dst)
 * (same multicast convention with flag_multicast).  `ctas` == 0 picks the default (32 CTAs of 512
 * threads); `ctas` < 0 moves the data with the copy engines instead (cudaMemcpyAsync per destination,
 * no SM touches the data) and publishes the flags from a one-CTA kernel behind them.
 * Asynchronous on `stream`. */
int b200_replicate_push(void* const* dst, int n_dst, int multicast, const void* src, size_t bytes,
                        uint32_t* const* flag_dst, int n_flag_dst, int flag_multicast,
                        uint32_t flag_value, int ctas, void* stream);

/* Same for a 2-D region: `rows` runs of `row_bytes` bytes (multiple of 16) with separate source and
 * destination pitches (bytes, multiples of 16) — a column panel of a row-major matrix. */
int b200_replicate_push_2d(void* const* dst, int n_dst, int multicast, const void* src, size_t rows,
                           size_t row_bytes, size_t src_pitch, size_t dst_pitch,
                           uint32_t* const* flag_dst, int n_flag_dst, int flag_multicast,
                           uint32_t flag_value, int ctas, void* stream);

/* Block `stream` (one resident warp, no host involvement) until flag[i*stride] has reached
 * `value` for every i < count (count <= 32) except i == skip (skip < 0: none).  Traps after ~30 s. */
int b200_flag_wait(const uint32_t* flag, uint32_t value, int count, int stride, int skip, void* stream);

/* Write `value` to *flag (usually a peer's flag word) after all prior work of `stream`. */
int b200_flag_signal(uint32_t* flag, uint32_t value, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_REPLICATE_H */
