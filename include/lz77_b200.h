/*
 * lz77_b200.h -- C ABI of the B200-native LZ77 codec (liblz77b200.so).
 *
 * Drop-in boundary for the hot path of cstdvd/lz77 (citations are file:line
 * into the reference tree):
 *
 *   reference interface                         replaced by
 *   ------------------------------------------  ---------------------------------
 *   encode(FILE*, bitFILE*, la, sb) lz77.h:14   lz77_gpu_encode[_device]
 *     lz77.c:51-140 driving tree.c:62-260
 *     (insert/find/delete/updateOffset) and
 *     writecode lz77.c:246-252 -> bitIO_write
 *     bitio.c:203-239
 *   decode(bitFILE*, FILE*)        lz77.h:15    lz77_gpu_decode[_device],
 *     lz77.c:148-197, readcode lz77.c:260-283       lz77_gpu_decode_size[_device]
 *     -> bitIO_read bitio.c:256-298
 *   bitof(n)                       bitio.h:25   lz77_bitof / lz77_token_bits
 *     bitio.c:41-43
 *   defaults LA 15 / SB 4095       lz77.c:21-22 LZ77_DEFAULT_LA / LZ77_DEFAULT_SB
 *   limits  -l 2..255, -s 0..65535 main.c:35-38 LZ77_MIN_LA .. LZ77_MAX_SB
 *
 * The byte stream is the reference's: a 4-byte header
 * (SB, LA as uint16 LE; lz77.c:74-75) followed by fixed-width LSB-first tokens
 * off:bitof(SB) | len:bitof(LA) | next:8 (lz77.c:249-251), zero padded to a
 * byte (bitio.c:180-182).  The reference decoder decodes what the encoder
 * here writes; the decoder here decodes what the reference encoder writes.
 *
 * Conventions: plain pointers and sizes; the caller owns every buffer; every
 * entry point returns 0 or a negative LZ77_E_* code and never calls exit();
 * there is NO CPU fallback -- without a CUDA device every compute entry point
 * returns LZ77_E_NODEVICE.  One host thread drives one GPU: lz77_gpu_init()
 * binds the calling thread to a device, and the state of every device is its
 * own, so several threads (or several processes) drive several GPUs side by
 * side.  Calls are synchronous (the result is complete on return).  Not
 * re-entrant on the same device from several threads.
 */
#ifndef LZ77_B200_H
#define LZ77_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define LZ77_DEFAULT_LA 15    /* lz77.c:21 */
#define LZ77_DEFAULT_SB 4095  /* lz77.c:22 */
#define LZ77_MIN_LA 2         /* main.c:35 */
#define LZ77_MAX_LA 255       /* main.c:36 */
#define LZ77_MIN_SB 0         /* main.c:37 (0 is accepted by the CLI; see B3) */
#define LZ77_MAX_SB 65535     /* main.c:38 */

#define LZ77_OK          0
#define LZ77_E_ARG      (-1)  /* bad parameter (size, sb/la out of range, NULL)   */
#define LZ77_E_SPACE    (-2)  /* output buffer too small                          */
#define LZ77_E_STREAM   (-3)  /* malformed stream (short header, sb/la of 0, an   */
                              /* offset that reaches before the start of output)  */
#define LZ77_E_NOMEM    (-4)  /* host or device allocation failed                 */
#define LZ77_E_NODEVICE (-5)  /* no CUDA device / library not initialised         */
#define LZ77_E_CUDA     (-6)  /* a CUDA call failed; see lz77_gpu_last_error()    */
#define LZ77_E_COMM     (-7)  /* NCCL failed / no communicator; lz77_gpu_last_error() */

/* ---- format arithmetic (host only, no device needed) -------------------- */

/* ceil(log2 n), the integer restatement of bitof(), bitio.c:41-43 */
int  lz77_bitof(int n);
/* bits per token: bitof(sb) + bitof(la) + 8, lz77.c:249-251 */
int  lz77_token_bits(int sb, int la);
/* worst case stream size: header + one token per input byte */
long lz77_gpu_encode_bound(long n_in, int sb, int la);
/* size of the independent blocks the encoder cuts the input into: no match
 * reaches across a multiple of this, and a token starts on every multiple */
long lz77_gpu_block_size(int sb);
/* bytes after which the greedy parse restarts inside a block, for these parameters
 * (-1: the defaults) and the current encoder options (1024 today; callers that restate
 * the encoder's specification ask instead of assuming) */
long lz77_gpu_segment_size(int sb, int la);

/* ---- lifetime ------------------------------------------------------------ */

int  lz77_gpu_device_count(void);
/* Bind the calling thread to one CUDA device (creates the device's stream and
 * scratch on first use).  Calling it again with another device re-binds the
 * thread; the state of the first device stays alive.  Threads that never
 * called it use the device of the most recent call. */
int  lz77_gpu_init(int device);
/* releases the state of every device; no other thread may be in the library */
void lz77_gpu_shutdown(void);
const char *lz77_gpu_strerror(int rc);
const char *lz77_gpu_last_error(void);

/* Run every later call on the caller's CUDA stream (a cudaStream_t passed as
 * void*) instead of the library's own; NULL restores the library's stream.
 * Calls stay synchronous: the stream is drained before they return. */
int  lz77_gpu_set_stream(void *cuda_stream);

/* The host entry points pipeline inputs larger than this many bytes in chunks
 * (H2D copy, kernels and D2H copy of successive chunks overlap; default 8 MiB,
 * 16 MiB for windows above 8191 bytes).
 * bytes <= 0 turns chunking off. */
void lz77_gpu_set_host_chunk(long bytes);

/* History mode of the encoder (default off).  Off: independent blocks -- no match
 * reaches across a multiple of lz77_gpu_block_size(), which is what lets the blocks of
 * one stream be decoded side by side and one stream be decoded by several GPUs.  On: the
 * match window slides across block seams exactly like the reference's (the last SB
 * bytes, wherever they are: lz77.c:101-105, tree.c:118-152), which recovers the
 * reference's compression ratio at large windows; such streams decode by pointer
 * jumping (like streams of the reference encoder) and cannot shard for decode.  Either
 * way the stream is the reference's format and decodes with the reference decoder. */
void lz77_gpu_set_history(int enabled);

/* Encoder, 24-bit tokens (the default parameters) and windows <= 8191 only (default
 * off).  On: the bit-packer (writecode lz77.c:246-252, bitIO_write bitio.c:203-239)
 * runs inside the search kernel -- tokens never travel through HBM unpacked, encode
 * DRAM traffic is N + C instead of N + 9 K + C and the scratch shrinks from 4 bytes per
 * input byte to a few MiB -- at the price of an ordered hand-over between tiles that
 * currently costs more time than the separate pack kernel saves.  The
 * stream is byte-identical either way. */
void lz77_gpu_set_fused_pack(int enabled);

/* Streams the reference encoder wrote (matches that leave their block) are
 * decoded by pointer jumping over pieces of this many output bytes (default
 * 64 MiB, 1 MiB .. 256 MiB; the scratch is 12 bytes per byte of a piece).
 * 0 restores the default. */
int lz77_gpu_set_jump_piece(long bytes);

/* pinned host memory for fast host<->device copies (optional) */
void *lz77_gpu_host_alloc(long n);
void  lz77_gpu_host_free(void *p);

/* ---- host-buffer entry points (what encode()/decode() wrappers call) ---- */

/* sb / la of -1 select the defaults, as in encode(), lz77.c:65-66 */
int lz77_gpu_encode(const unsigned char *in, long n_in, int sb, int la,
                    unsigned char *out, long out_cap, long *n_out);
int lz77_gpu_decode_size(const unsigned char *in, long n_in, long *n_out);
int lz77_gpu_decode(const unsigned char *in, long n_in,
                    unsigned char *out, long out_cap, long *n_out);

/* ---- device-buffer entry points ------------------------------------------
 * Pointers are device memory on the bound device, 16-byte aligned, with a
 * capacity that is a multiple of 16 bytes (the kernels move 128-bit words).
 * The caller must have finished producing d_in before the call. */
int lz77_gpu_encode_device(const void *d_in, long n_in, int sb, int la,
                           void *d_out, long out_cap,
                           long *n_out, long *n_tokens);
int lz77_gpu_decode_size_device(const void *d_in, long n_in, long *n_out);
int lz77_gpu_decode_device(const void *d_in, long n_in,
                           void *d_out, long out_cap, long *n_out);

/* ---- token-array helpers for sharding one stream across GPUs --------------
 * Tokens are fixed width, so a stream splits at any token without parsing
 * (the decoder's counterpart of lz77.c:260-283 reading one token at bit
 * 32 + k*T).  slice_tokens writes a standalone stream: the header of d_in and
 * tokens [tok_lo, tok_hi) (out_cap a multiple of 4 >= 4 + ceil((hi-lo)*T/8)
 * rounded up to 4).  token_at returns the token that holds decoded byte `pos`
 * (the token that starts there when one does) and the decoded position of its
 * first byte; pos == decoded size gives (token count, decoded size). */
int lz77_gpu_slice_tokens_device(const void *d_in, long n_in, long tok_lo, long tok_hi,
                                 void *d_out, long out_cap, long *n_out);
int lz77_gpu_token_at_device(const void *d_in, long n_in, long pos,
                             long *tok, long *tok_pos);

/* ---- several GPUs: one input, one stream ----------------------------------
 * The encoder's blocks are independent (no reference counterpart: the reference
 * is one sequential loop, lz77.c:89-135), so a run of whole blocks can be encoded
 * on any GPU and the token payloads of consecutive runs concatenate into exactly
 * the stream one GPU writes for the whole input: one header (lz77.c:74-75), then
 * fixed-width tokens back to back (lz77.c:246-252).  Ranks are one per GPU --
 * processes (torchrun, MPI) or threads of one process -- joined by an NCCL
 * communicator the library owns: rank 0 obtains an id, every rank passes the
 * same id to lz77_comm_init() after lz77_gpu_init(its device).
 *
 *   encode_sharded  scatter of block runs from root (grouped ncclSend/ncclRecv
 *                   over NVLink) -> every rank encodes its run -> all-gather of
 *                   the token counts -> the payloads travel to root, straight to
 *                   their byte offset when T is a multiple of 8, through a
 *                   device-side bit shift (<= 7-bit seam merge) otherwise
 *   decode_sharded  tokens are fixed width, so root cuts the token array evenly
 *                   (plus one block of margin) without parsing -> every rank sums
 *                   len+1 over its tokens -> all-gather -> every rank finds the
 *                   token that starts the first block at or after its position ->
 *                   all-gather of these split points -> every rank decodes its
 *                   run of whole blocks -> the plaintext travels to root.  Only
 *                   streams of the block encoder shard; a stream of the
 *                   reference encoder is LZ77_E_STREAM on every rank (replicas
 *                   only: each GPU would need its predecessor's last SB bytes).
 *
 * Both are collective: every rank of the communicator calls them; d_in / d_out /
 * n_in / sb / la / out_cap count on root only (device pointers, 16-byte aligned).
 * Every rank returns the same code; n_out / n_tokens are set on every rank. */
#define LZ77_COMM_ID_BYTES 128
/* byte range [*lo, *hi) of rank `rank` of `world`: contiguous runs of whole
 * blocks, as even as whole blocks allow (host only, no device needed) */
int lz77_shard_range(long n_bytes, int world, long block, int rank, long *lo, long *hi);
int lz77_comm_get_unique_id(void *id /* LZ77_COMM_ID_BYTES */);
int lz77_comm_init(const void *id, int rank, int world);
void lz77_comm_destroy(void);
int lz77_gpu_encode_sharded_device(const void *d_in, long n_in, int sb, int la,
                                   void *d_out, long out_cap,
                                   long *n_out, long *n_tokens, int root);
int lz77_gpu_decode_sharded_device(const void *d_in, long n_in,
                                   void *d_out, long out_cap, long *n_out, int root);
/* what the last sharded call of this rank moved through the communicator */
struct lz77_comm_stats {
    long  sent_bytes, recv_bytes;  /* payload bytes this rank sent / received    */
    float scatter_ms;              /* input (encode) / token slices (decode)     */
    float compute_ms;              /* this rank's kernels                        */
    float gather_ms;               /* payloads (encode) / plaintext (decode)     */
    float total_ms;
    int   collectives;             /* NCCL calls (groups count once)             */
};
int lz77_comm_last_stats(struct lz77_comm_stats *s);

/* Single process, n_gpus devices (what the command-line program's -G uses): the
 * library keeps one worker thread per device, each a rank of its own
 * communicator.  Host buffers; root = device 0 copies them in and out. */
int  lz77_mgpu_init(int n_gpus);
void lz77_mgpu_shutdown(void);
int  lz77_mgpu_encode(const unsigned char *in, long n_in, int sb, int la,
                      unsigned char *out, long out_cap, long *n_out);
int  lz77_mgpu_decode(const unsigned char *in, long n_in,
                      unsigned char *out, long out_cap, long *n_out);

/* ---- measurement ---------------------------------------------------------
 * Device time (CUDA events on the library's stream) of each kernel of the
 * last encode / decode call, in milliseconds, and how many kernels ran. */
struct lz77_timing {
    float enc_search_ms;   /* longest-match search + greedy parse           */
    float enc_scan_ms;     /* token count prefix sums                       */
    float enc_pack_ms;     /* warp-cooperative bit-packer                   */
    float dec_scan_ms;     /* token length scan + tile table                */
    float dec_copy_ms;     /* match-copy / literal tile decode              */
    float h2d_ms, d2h_ms;  /* host entry points only                        */
    int   launches;        /* kernels launched by the last call             */
    long  n_tokens;        /* tokens written / read by the last call        */
};
int lz77_gpu_last_timing(struct lz77_timing *t);
/* per-kernel event timing costs a few synchronisations; 0 turns it off */
void lz77_gpu_set_timing(int enabled);

#ifdef __cplusplus
}
#endif
#endif /* LZ77_B200_H */
