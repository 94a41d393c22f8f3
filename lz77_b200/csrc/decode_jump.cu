// decode_jump.cu -- decoder for streams whose matches are not confined to blocks
// (what the reference encoder writes: lz77.c:51-140 parses the whole file through
// one sliding window, so a match may reach back SB bytes from anywhere).
//
// The tile decoder of decode.cu needs the tail of tile j-1 before the head of tile
// j, which turns such a stream into a chain of tiles.  Here the copy semantics of
// lz77.c:178-194 are resolved for all bytes at once by pointer jumping:
//
//   build    every output byte i of a range gets one 32-bit state word S[i]:
//              FINAL | value      a literal, or a match byte whose source lies
//                                 before the range (already decoded, read from HBM)
//              i - off            the (range-relative) index of its source byte
//            (a self-overlapping match needs no special case: byte i of the match
//            simply points at byte i - off of the same match)
//   jump     S[i] <- S[S[i]] for every word that is not final, kHops times per
//            pass.  Copying the source's word either copies its final value or its
//            pointer, i.e. halves the distance to the nearest final ancestor, so
//            ceil(log_{kHops+1}(depth)) + 1 passes resolve the deepest chain.  Words
//            are 32-bit and every thread writes only its own, so a pass may read a
//            word before or after its owner updated it: both are ancestors.
//            Every pass counts the words it left open; the following passes return
//            at once when that count is zero, so the host queues the worst-case
//            number of passes without ever reading the count.
//   extract  out[i] = S[i] & 0xff, 16 bytes per thread.
//
// A range is processed in pieces of kJumpPiece output bytes so that S (4 bytes per
// output byte) stays L2-resident between the passes.
#include "kernels.cuh"

namespace lz77 {

#ifndef LZ77_JUMP_PIECE_MIB
#define LZ77_JUMP_PIECE_MIB 16
#endif
#ifndef LZ77_JUMP_HOPS
#define LZ77_JUMP_HOPS 2
#endif
constexpr long long kJumpPiece = (long long)LZ77_JUMP_PIECE_MIB << 20;
constexpr int kJumpHops = LZ77_JUMP_HOPS;
constexpr uint32_t kFinal = 0x80000000u;
constexpr int kJumpThreads = 256;
constexpr int kMaxJumpPasses = 40;

size_t decode_jump_scratch_bytes()
{
    return (size_t)kJumpPiece * 4 + 64 + 4096;  // S, padding, pass counters
}

// ---- build -------------------------------------------------------------------

__global__ void __launch_bounds__(kJumpThreads)
lz77_jump_build_kernel(const uint32_t *__restrict__ words, long long n_words, long long n_tokens,
                       Params P, int tile_shift, const long long *__restrict__ tile_tok,
                       const long long *__restrict__ tile_pos,
                       const uint32_t *__restrict__ group_pos, long long lo, long long hi,
                       bool to_end, const uint8_t *out, uint32_t *S, unsigned int *open,
                       DecodeInfo *info)
{
    __shared__ unsigned int s_open;
    if (threadIdx.x == 0) s_open = 0u;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const uint32_t off_mask = (1u << P.ob) - 1u;
    const uint32_t len_mask = (1u << P.lb) - 1u;
    const int lit_shift = P.ob + P.lb;
    const int n = (int)(hi - lo);

    const long long k0 = tile_tok[lo >> tile_shift];
    long long k_end = n_tokens;
    if (!to_end) {
        const long long j = hi >> tile_shift;  // hi is a tile boundary here
        k_end = tile_tok[j] + (tile_pos[j] < hi ? 1 : 0);
    }
    const long long g_first = k0 >> 5;
    const long long n_groups = ((k_end + 31) >> 5) - g_first;
    const uint32_t lo32 = (uint32_t)lo;

    // the words behind the range up to the next multiple of 16 are final zeros
    if (blockIdx.x == 0 && threadIdx.x < 16) {
        const int i = n + threadIdx.x;
        if (i < ((n + 15) & ~15)) S[i] = kFinal;
    }

    const long long warp_id = (long long)blockIdx.x * (kJumpThreads / 32) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (kJumpThreads / 32);
    bool any_open = false;
    for (long long gi = warp_id; gi < n_groups; gi += n_warps) {
        const long long k = ((g_first + gi) << 5) + lane;
        uint32_t tok = 0;
        if (k < k_end) tok = load_bits32(words, n_words, kHeaderBits + k * P.tbits);
        const int off = (int)(tok & off_mask);
        const int len = (int)((tok >> P.ob) & len_mask);
        const int l1 = k < k_end ? len + 1 : 0;  // lanes before k0 still count in the sum
        int inc = l1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const int pos_rel = (int)(__ldg(group_pos + g_first + gi) - lo32) + inc - l1;
        if (k < k0 || k >= k_end) continue;

        // literal, lz77.c:189-194
        const int d_lit = pos_rel + len;
        if (d_lit >= 0 && d_lit < n) S[d_lit] = kFinal | ((tok >> lit_shift) & 0xffu);
        if (len == 0) continue;
        // match bytes, lz77.c:178-188
        const int i_lo = max(pos_rel, 0), i_hi = min(pos_rel + len, n);
        if (off == 0 || (long long)off > lo + pos_rel) {  // source before the start of the output
            info->error = 1;
            for (int i = i_lo; i < i_hi; i++) S[i] = kFinal;
            continue;
        }
        for (int i = i_lo; i < i_hi; i++) {
            const int s = i - off;
            if (s < 0) {
                // before the range: final in HBM (also the earlier bytes of this very match
                // when it straddles the range start)
                S[i] = kFinal | (uint32_t)__ldcg(out + (lo + s));
            } else {
                S[i] = (uint32_t)s;
                any_open = true;
            }
        }
    }
    if (__any_sync(0xffffffffu, any_open) && lane == 0) s_open = 1u;
    __syncthreads();
    if (threadIdx.x == 0 && s_open) atomicOr(open, 1u);
}

// ---- jump ----------------------------------------------------------------------

__global__ void __launch_bounds__(kJumpThreads)
lz77_jump_pass_kernel(uint32_t *S, int n_vec, const unsigned int *open_in, unsigned int *open_out)
{
    if (*open_in == 0u) return;  // resolved by an earlier pass
    bool left = false;
    uint4 *S4 = reinterpret_cast<uint4 *>(S);
    for (int v = blockIdx.x * kJumpThreads + threadIdx.x; v < n_vec; v += gridDim.x * kJumpThreads) {
        const uint4 w = __ldcg(S4 + v);
        if ((w.x & w.y & w.z & w.w) & kFinal) continue;
        uint32_t a[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int h = 0; h < kJumpHops; h++) {
            uint32_t t[4];
#pragma unroll
            for (int c = 0; c < 4; c++) t[c] = (a[c] & kFinal) ? a[c] : __ldcg(S + a[c]);
#pragma unroll
            for (int c = 0; c < 4; c++) a[c] = t[c];
        }
        left |= ((a[0] & a[1] & a[2] & a[3]) & kFinal) == 0u;
        __stcg(S4 + v, make_uint4(a[0], a[1], a[2], a[3]));
    }
    if (__any_sync(0xffffffffu, left) && (threadIdx.x & 31) == 0) atomicOr(open_out, 1u);
}

// ---- extract -------------------------------------------------------------------

__device__ __forceinline__ uint32_t pack4(const uint4 &w)
{
    return (w.x & 0xffu) | ((w.y & 0xffu) << 8) | ((w.z & 0xffu) << 16) | ((w.w & 0xffu) << 24);
}

__global__ void __launch_bounds__(kJumpThreads)
lz77_jump_extract_kernel(const uint32_t *S, int n, uint8_t *out /* + lo, 16-byte aligned */,
                         const unsigned int *open_last, DecodeInfo *info)
{
    if (blockIdx.x == 0 && threadIdx.x == 0 && *open_last != 0u) info->error = 2;  // cannot happen
    const uint4 *S4 = reinterpret_cast<const uint4 *>(S);
    const int n16 = n >> 4;
    for (int v = blockIdx.x * kJumpThreads + threadIdx.x; v < n16; v += gridDim.x * kJumpThreads) {
        uint4 o;
        o.x = pack4(__ldcg(S4 + 4 * v));
        o.y = pack4(__ldcg(S4 + 4 * v + 1));
        o.z = pack4(__ldcg(S4 + 4 * v + 2));
        o.w = pack4(__ldcg(S4 + 4 * v + 3));
        reinterpret_cast<uint4 *>(out)[v] = o;
    }
    if (blockIdx.x == 0) {
        for (int i = (n16 << 4) + threadIdx.x; i < n; i += kJumpThreads)
            out[i] = (uint8_t)(__ldcg(S + i) & 0xffu);
    }
}

// ---- launcher ------------------------------------------------------------------

// Decodes output bytes [out_lo, out_hi) of a stream whose token scan has covered
// them; everything before out_lo is final in d_out (stream order).  out_lo is a
// tile boundary; out_hi is a tile boundary, or the end of the output with
// to_end = true.
cudaError_t launch_decode_jump_range(const uint32_t *d_in_words, long long n_in_bytes,
                                     long long n_tokens, long long out_lo, long long out_hi,
                                     bool to_end, const Params &P, void *scratch,
                                     void *jump_scratch, uint8_t *d_out, cudaStream_t st)
{
    if (out_hi <= out_lo) return cudaSuccess;
    const DecodeTables t = decode_tables(scratch, n_tokens, P);
    uint32_t *S = reinterpret_cast<uint32_t *>(jump_scratch);
    unsigned int *open = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(jump_scratch) +
                                                          (size_t)kJumpPiece * 4 + 64);
    const long long n_words = (n_in_bytes + 3) / 4;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 8;

    for (long long lo = out_lo; lo < out_hi; lo += kJumpPiece) {
        const long long hi = lo + kJumpPiece < out_hi ? lo + kJumpPiece : out_hi;
        const bool piece_to_end = to_end && hi == out_hi;
        const int n = (int)(hi - lo);
        // depth <= n, every pass divides it by kJumpHops + 1, one more pass finalises
        int passes = 2;
        for (long long d = 1; d < n; d *= (kJumpHops + 1)) passes++;
        if (passes > kMaxJumpPasses) passes = kMaxJumpPasses;
        cudaError_t rc = cudaMemsetAsync(open, 0, (kMaxJumpPasses + 2) * sizeof(unsigned int), st);
        if (rc != cudaSuccess) return rc;
        lz77_jump_build_kernel<<<grid, kJumpThreads, 0, st>>>(
            d_in_words, n_words, n_tokens, P, P.tile_shift, t.tile_tok, t.tile_pos, t.group_pos, lo,
            hi, piece_to_end, d_out, S, open, t.info);
        const int n_vec = (n + 3) >> 2;
        for (int p = 0; p < passes; p++)
            lz77_jump_pass_kernel<<<grid, kJumpThreads, 0, st>>>(S, n_vec, open + p, open + p + 1);
        lz77_jump_extract_kernel<<<grid, kJumpThreads, 0, st>>>(S, n, d_out + lo, open + passes,
                                                                t.info);
    }
    return cudaGetLastError();
}

}  // namespace lz77
