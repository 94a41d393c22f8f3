// kernels.cuh -- launch interface between the C ABI (capi.cu) and the kernels.
#pragma once

#include "common.cuh"

namespace lz77 {

struct StageEvents {
    cudaEvent_t e[4];
};

// ---- encoder (encode.cu) ---------------------------------------------------
size_t encode_scratch_bytes(long long n_in, const Params &P);
int encode_launch_count(long long n_in, const Params &P);
// d_out_words must hold 4 + ceil(n_in * T / 8) bytes rounded up to 16.
// *d_total_tokens receives a device pointer (inside scratch) to the token count.
// `pre`: valid input bytes in front of d_in (history mode only; 0 otherwise)
cudaError_t launch_encode(const uint8_t *d_in, long long n_in, long long pre, const Params &P,
                          void *scratch, uint32_t *d_out_words,
                          unsigned long long **d_total_tokens, cudaStream_t st, StageEvents *ev);

// chunked encoding (the host entry point overlaps copies with kernels)
struct EncodePlan {
    uint32_t *tok_tmp;
    uint32_t *seg_ntok;
    unsigned long long *prefix;
    unsigned long long *partial;
    unsigned long long *total;  // running token count (device)
    void *big;                  // large-window bucket tables (search_bigwin.cu)
    void *fused;                // fused search + pack path: look-back words, tickets, spill
    long long n_total;          // input bytes of the whole call
};
EncodePlan encode_plan(void *scratch, long long n_in_total, const Params &P);
long long encode_chunk_granule();
cudaError_t launch_encode_chunk(const uint8_t *d_in_base, long long pre_base, long long lo,
                                long long n_chunk, bool first, const Params &P,
                                const EncodePlan &pl,
                                uint32_t *d_out_words, cudaStream_t st, StageEvents *ev,
                                int phase, unsigned long long *host_total, int slot);

// bucketed longest-match search + greedy parse (search_bucket.cu)
// `pre`: valid input bytes in front of d_in (history mode reaches back into them)
cudaError_t launch_parse_bucket(const uint8_t *d_in, long long n_in, long long pre,
                                const Params &P, uint32_t *tok_tmp, uint32_t *seg_ntok,
                                cudaStream_t st);

// bytes after which the greedy parse restarts (part of the stream's specification)
int parse_segment_bytes(int window, int la, bool fused_pack);

// 24-bit tokens and a window <= 8191: search + parse + pack in one persistent kernel
// (search_bucket.cu); `slot` < 64 names the ticket of this launch (launches of one call
// that may run side by side need different ones)
bool parse_bucket_fused(const Params &P);
size_t parse_bucket_fused_scratch(long long n_total);
cudaError_t launch_parse_bucket_fused_reset(long long n_total, void *scratch, cudaStream_t st);
cudaError_t launch_parse_bucket_fused(const uint8_t *d_in, long long lo, long long n_in,
                                      long long n_total, long long pre, bool first, bool reset,
                                      int slot, const Params &P, void *scratch, uint8_t *d_out,
                                      unsigned long long *total, unsigned long long *host_total,
                                      cudaStream_t st);

// large windows: block-level buckets (search_bigwin.cu)
size_t bigwin_scratch_bytes(long long n_in, const Params &P);
cudaError_t launch_parse_bigwin(const uint8_t *d_in, long long n_in, long long pre,
                                const Params &P, void *scratch, uint32_t *tok_tmp,
                                uint32_t *seg_ntok, cudaStream_t st);

void bigwin_release(int device);  // side stream + events of the large-window encoder

// ---- decoder (decode.cu) ---------------------------------------------------
struct DecodeInfo {            // lives in device scratch, copied back by the C ABI
    unsigned long long n_out;  // decoded size
    unsigned int error;        // != 0: malformed stream (bad offset)
    unsigned int cross_block;  // != 0: a match source lies before its block (not a stream of
                               // the block-parallel encoder): tiles must run in order
};

// the tables pass 1 leaves in the decode scratch area
struct DecodeTables {
    const long long *tile_tok;  // token containing the first byte of tile j
    const long long *tile_pos;  // output position of that token
    const uint32_t *group_pos;  // low 32 bits of the output position of token 32g
    DecodeInfo *info;
};
DecodeTables decode_tables(void *scratch, long long n_tokens, const Params &P);

int decode_tile_bytes(const Params &P);
size_t decode_scratch_bytes(long long n_tokens, const Params &P);
int decode_launch_count(bool with_copy);
// pass 1: token lengths -> decoded size + tile table
cudaError_t launch_decode_scan(const uint32_t *d_in_words, long long n_in_bytes,
                               long long n_tokens, const Params &P, void *scratch,
                               DecodeInfo **d_info, cudaStream_t st);
// ranged variants for the chunked host path (see decode.cu)
long long decode_scan_granule();
cudaError_t launch_decode_scan_range(const uint32_t *d_in_words, long long n_in_bytes,
                                     long long n_tokens, long long tok_begin, long long tok_end,
                                     const Params &P, void *scratch, DecodeInfo **d_info,
                                     cudaStream_t st, unsigned long long *host_n_out);
cudaError_t launch_decode_tiles_range(const uint32_t *d_in_words, long long n_in_bytes,
                                      long long n_tokens, long long tile_begin,
                                      long long tile_end, bool last, long long n_out,
                                      int launch_idx, int pair_mode, const Params &P,
                                      void *scratch, uint8_t *d_out, cudaStream_t st);
// token-array helpers (sharding one stream across GPUs)
cudaError_t launch_slice_tokens(const uint32_t *d_in_words, long long n_in_bytes,
                                long long tok_lo, long long tok_hi, const Params &P,
                                uint32_t *d_out_words, long long out_words, cudaStream_t st);
cudaError_t launch_token_at(const uint32_t *d_in_words, long long n_in_bytes, long long n_tokens,
                            long long pos, const Params &P, void *scratch, long long *d_result,
                            cudaStream_t st);
// pass 2 (needs the decoded size pass 1 produced): tile decode, or pointer jumping
// (decode_jump.cu; jump_scratch of decode_jump_scratch_bytes()) when cross_block
cudaError_t launch_decode_copy(const uint32_t *d_in_words, long long n_in_bytes,
                               long long n_tokens, long long n_out, bool cross_block,
                               const Params &P, void *scratch, void *jump_scratch,
                               long long jump_piece, uint8_t *d_out, cudaStream_t st);
// output bytes per round (override: bytes per piece, 0 = the default)
long long decode_jump_piece(long long n_out_max, const Params &P, long long override_bytes);
size_t decode_jump_scratch_bytes(long long jump_piece);
cudaError_t launch_decode_jump_range(const uint32_t *d_in_words, long long n_in_bytes,
                                     long long n_tokens, long long out_lo, long long out_hi,
                                     bool to_end, const Params &P, void *scratch,
                                     void *jump_scratch, long long jump_piece, uint8_t *d_out,
                                     cudaStream_t st);

}  // namespace lz77
