/*
 * lz77_oracle.h -- CPU restatement of the cstdvd/lz77 codec hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the reported CPU baseline.  The product path (lz77_b200/) never links,
 * imports or executes it and fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED.  lz77o_ref_encode() is byte-identical to the
 * compiled reference (oracle/_ref/lz77, built from /root/reference by
 * oracle/Makefile) on every known-answer vector of SURVEY.md Appendix C
 * (tests/golden/) and on seeded random/text inputs (tests/test_oracle.py).
 *
 * All entry points work on whole in-memory buffers; the reference's FILE*
 * streaming is emulated exactly where it affects the output (window scroll
 * points, fread/feof semantics).
 */
#ifndef LZ77_ORACLE_H
#define LZ77_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LZ77O_DEFAULT_LA 15   /* reference lz77.c:21 */
#define LZ77O_DEFAULT_SB 4095 /* reference lz77.c:22 */

/* error codes (negative returns) */
#define LZ77O_E_ARG      (-1) /* bad parameter                              */
#define LZ77O_E_SPACE    (-2) /* output buffer too small                    */
#define LZ77O_E_STREAM   (-3) /* malformed stream (bad header, bad offset)  */
#define LZ77O_E_NOMEM    (-4)

/* ceil(log2 n) for n >= 1; restates bitio.c:41-43 in integers. */
int lz77o_bitof(int n);

/* bits per token for a (sb, la) pair: bitof(sb)+bitof(la)+8 (lz77.c:249-251) */
int lz77o_token_bits(int sb, int la);

/* worst case output size: every input byte its own token + 4-byte header */
long lz77o_encode_bound(long n_in, int sb, int la);

/*
 * Faithful restatement of the reference encoder (lz77.c:51-140 driving
 * tree.c:62-243 and bitio.c:203-239): same window scrolling, same BST
 * insert/find/delete, same tie-breaks => byte-identical output.
 * sb / la of -1 select the defaults.  Returns bytes written or <0.
 */
long lz77o_ref_encode(const uint8_t *in, long n_in, int sb, int la,
                      uint8_t *out, long out_cap);

/*
 * Restatement of the reference decoder (lz77.c:148-197, 260-283,
 * bitio.c:256-298).  out may be NULL to only compute the decoded size.
 * Returns decoded bytes or <0.  Unlike the reference it reports
 * LZ77O_E_STREAM instead of reading out of bounds on a bad offset.
 */
long lz77o_decode(const uint8_t *in, long n_in, uint8_t *out, long out_cap);

/*
 * Specification of the block-parallel encoder the GPU implements: the input
 * is cut into independent blocks of `block` bytes; inside a block each token
 * is the longest match (<= min(la, bytes left in block) - 1) against the last
 * min(sb, 2^bitof(sb) - 1, position in block) bytes, farthest offset winning
 * ties, followed by a literal -- i.e. the greedy parse tree.c:118-152 /
 * lz77.c:87-135 produce, restarted at every block.  block <= 0 means one
 * block (the whole input).  n_tokens (may be NULL) receives the token count.
 */
long lz77o_blocked_encode(const uint8_t *in, long n_in, int sb, int la,
                          long block, uint8_t *out, long out_cap,
                          long *n_tokens);

/*
 * The exact stream the GPU encoder writes: as lz77o_blocked_encode, and in
 * addition the greedy parse restarts every `segment` bytes inside a block (a
 * token never runs past the end of its segment; the match WINDOW still spans
 * the whole block).  segment must divide block; segment <= 0 means no restarts.
 * The GPU encoder uses block = lz77_gpu_block_size(sb), segment =
 * lz77_gpu_segment_size(); parity with it is byte-for-byte.  block <= 0 with
 * segment > 0 is the GPU encoder's history mode (lz77_gpu_set_history): the window
 * slides over the whole input like the reference's (lz77.c:101-105), only the parse
 * restarts remain.
 */
long lz77o_segmented_encode(const uint8_t *in, long n_in, int sb, int la,
                            long block, long segment, uint8_t *out,
                            long out_cap, long *n_tokens);

/*
 * Unpack a stream into token arrays (each of capacity cap, any may be NULL).
 * Returns the token count (which may exceed cap; only cap are stored) or <0.
 */
long lz77o_unpack_tokens(const uint8_t *in, long n_in, int *sb, int *la,
                         int32_t *off, int32_t *len, uint8_t *next, long cap);

#ifdef __cplusplus
}
#endif
#endif /* LZ77_ORACLE_H */
