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
//   jump     S[i] <- S[S[i]] for every word that is not final, h times per pass.
//            Copying the source's word either copies its final value or its
//            pointer, i.e. halves the distance to the nearest final ancestor, so
//            ceil(log_{h+1}(depth)) + 1 passes resolve the deepest chain.  Words
//            are 32-bit and every thread writes only its own, so a pass may read a
//            word before or after its owner updated it: both are ancestors.
//            The first pass sweeps all of S and lists the words it left open; every
//            later pass works through the list of its predecessor (cost proportional
//            to the open words) and writes the next list.  A pass whose input list is
//            empty returns at once, so the host queues the worst-case number of
//            passes without ever reading a count.
//   extract  out[i] = S[i] & 0xff, 16 bytes per thread.
//
// A range is processed in pieces of at most kJumpPiece output bytes, which bounds the
// scratch area (S and two index lists: 12 bytes per output byte of a piece) and the
// pointer width.  Measured on B200: larger pieces and more hops per pass are faster
// (fewer launches; S does not stay L2-resident even at 16 MiB pieces): 16 / 32 / 64 MiB
// pieces decode text at 60 / 67 / 75 GB/s.
#include "kernels.cuh"

namespace lz77 {

#ifndef LZ77_JUMP_PIECE_MIB
#define LZ77_JUMP_PIECE_MIB 64
#endif
#ifndef LZ77_JUMP_HOPS
#define LZ77_JUMP_HOPS 4
#endif
constexpr long long kJumpPiece = (long long)LZ77_JUMP_PIECE_MIB << 20;
constexpr int kJumpHops = LZ77_JUMP_HOPS;
#ifndef LZ77_JUMP_LATE_HOPS
#define LZ77_JUMP_LATE_HOPS 8
#endif
constexpr int kJumpLateHops = LZ77_JUMP_LATE_HOPS;
constexpr uint32_t kFinal = 0x80000000u;
constexpr int kJumpThreads = 256;
constexpr int kMaxJumpPasses = 40;

// Output bytes per piece for a stream that decodes to at most n_out_max bytes: the
// full piece, or the whole (tile-rounded) output when that is smaller.
long long decode_jump_piece(long long n_out_max, const Params &P, long long override_bytes)
{
    const long long tile = 1LL << P.tile_shift;
    const long long cap = override_bytes > 0 ? override_bytes : kJumpPiece;  // lz77_gpu_set_jump_piece()
    const long long whole = (n_out_max + tile - 1) / tile * tile;
    return whole < cap ? (whole > tile ? whole : tile) : cap;
}

// scratch layout: S | list A | list B | list lengths
static size_t jump_array_bytes(long long piece) { return (size_t)piece * 4 + 64; }

size_t decode_jump_scratch_bytes(long long piece) { return 3 * jump_array_bytes(piece) + 4096; }

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
    // per warp: one bit per output byte of the current group, set where a token starts
    extern __shared__ uint32_t s_heads_all[];
    const int head_words = 1 << P.lb;  // 32 tokens of at most 2^lb bytes
    uint32_t *heads = s_heads_all + (threadIdx.x >> 5) * head_words;
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
        const int l1 = k < k_end ? (int)((tok >> P.ob) & len_mask) + 1 : 0;  // lanes before k0 still count in the sum
        int inc = l1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        const int p0 = (int)(__ldg(group_pos + g_first + gi) - lo32);  // of token 32 * group
        const long long abs0 = lo + p0;
        __syncwarp();
        for (int w = lane; w * 32 < total; w += 32) heads[w] = 0u;
        __syncwarp();
        if (l1 > 0) atomicOr(&heads[(inc - l1) >> 5], 1u << ((inc - l1) & 31));
        __syncwarp();
        // one lane per output byte of the group (coalesced stores); the tokens before k0
        // and from k_end on lie outside [0, n) entirely
        int before = 0;  // tokens that start before this row of 32 bytes
        for (int b0 = 0; b0 < total; b0 += 32) {
            const int b = b0 + lane;
            const uint32_t hw = heads[b0 >> 5];
            // the token that holds byte b: the last one that starts at or before b
            const int t = before + __popc(hw & (0xffffffffu >> (31 - lane))) - 1;
            before += __popc(hw);
            const uint32_t tk = __shfl_sync(0xffffffffu, tok, t & 31);
            const int end_t = __shfl_sync(0xffffffffu, inc, t & 31);
            const int i = p0 + b;
            if (b >= total || i < 0 || i >= n) continue;
            const int off_t = (int)(tk & off_mask);
            const int len_t = (int)((tk >> P.ob) & len_mask);
            const int start_t = end_t - (len_t + 1);
            if (b - start_t == len_t) {  // literal, lz77.c:189-194
                S[i] = kFinal | ((tk >> lit_shift) & 0xffu);
            } else if (off_t == 0 || (long long)off_t > abs0 + start_t) {
                info->error = 1;  // source before the start of the output
                S[i] = kFinal;
            } else {  // match byte, lz77.c:178-188
                const int src = i - off_t;
                if (src < 0) {
                    // before the range: final in HBM (also the earlier bytes of this very
                    // match when it straddles the range start)
                    S[i] = kFinal | (uint32_t)__ldcg(out + (lo + src));
                } else {
                    S[i] = (uint32_t)src;
                    any_open = true;
                }
            }
        }
    }
    if (__any_sync(0xffffffffu, any_open) && lane == 0) s_open = 1u;
    __syncthreads();
    if (threadIdx.x == 0 && s_open) atomicOr(open, 1u);
}

// ---- jump ----------------------------------------------------------------------

// Each thread owns kJumpPer words, kJumpThreads apart, so the loads, the gathers of
// consecutive match bytes and the stores are all coalesced.
#ifndef LZ77_JUMP_PER
#define LZ77_JUMP_PER 4
#endif
constexpr int kJumpPer = LZ77_JUMP_PER;
constexpr int kJumpSpan = kJumpThreads * kJumpPer;

// Appends the indices whose word is still open to the list: one global atomic per
// block and round (a single counter cannot take one atomic per warp).  `round`
// alternates the shared buffers so that no barrier is needed behind the reads.
struct AppendShared {
    unsigned int warp_total[2][kJumpThreads / 32];
    unsigned int base[2];
};

__device__ __forceinline__ void block_list_append(const bool (&open)[kJumpPer],
                                                  const uint32_t (&idx)[kJumpPer], uint32_t *list,
                                                  unsigned int *count, AppendShared &sh, int round)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = round & 1;
    unsigned m[kJumpPer];
    unsigned wt = 0;
#pragma unroll
    for (int c = 0; c < kJumpPer; c++) {
        m[c] = __ballot_sync(0xffffffffu, open[c]);
        wt += __popc(m[c]);
    }
    if (lane == 0) sh.warp_total[r][warp] = wt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
#pragma unroll
        for (int w = 0; w < kJumpThreads / 32; w++) tot += sh.warp_total[r][w];
        sh.base[r] = tot ? atomicAdd(count, tot) : 0u;
    }
    __syncthreads();
    unsigned at = sh.base[r];
    for (int w = 0; w < warp; w++) at += sh.warp_total[r][w];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int c = 0; c < kJumpPer; c++) {
        if (open[c]) list[at + __popc(m[c] & lt)] = idx[c];
        at += __popc(m[c]);
    }
}

// first pass: all words of S
template <int kHops>
__global__ void __launch_bounds__(kJumpThreads)
lz77_jump_sweep_kernel(uint32_t *S, int n, const unsigned int *any_open, uint32_t *list_out,
                       unsigned int *count_out)
{
    __shared__ AppendShared sh;
    if (*any_open == 0u) return;  // literals only
    int round = 0;
    for (long long base = (long long)blockIdx.x * kJumpSpan; base < n;
         base += (long long)gridDim.x * kJumpSpan, round++) {
        uint32_t a[kJumpPer];
        bool was_open[kJumpPer];
#pragma unroll
        for (int c = 0; c < kJumpPer; c++) {
            const long long i = base + c * kJumpThreads + threadIdx.x;
            a[c] = i < n ? __ldcg(S + i) : kFinal;
            was_open[c] = (a[c] & kFinal) == 0u;
        }
#pragma unroll
        for (int h = 0; h < kHops; h++) {
#pragma unroll
            for (int c = 0; c < kJumpPer; c++)
                if ((a[c] & kFinal) == 0u) a[c] = __ldcg(S + a[c]);
        }
        uint32_t idx[kJumpPer];
        bool open[kJumpPer];
#pragma unroll
        for (int c = 0; c < kJumpPer; c++) {
            idx[c] = (uint32_t)(base + c * kJumpThreads + threadIdx.x);
            if (was_open[c]) __stcg(S + idx[c], a[c]);
            open[c] = (a[c] & kFinal) == 0u;
        }
        block_list_append(open, idx, list_out, count_out, sh, round);
    }
}

// later passes: the words the previous pass left open
template <int kHops>
__global__ void __launch_bounds__(kJumpThreads)
lz77_jump_list_kernel(uint32_t *S, const uint32_t *list_in, const unsigned int *count_in,
                      uint32_t *list_out, unsigned int *count_out)
{
    __shared__ AppendShared sh;
    const unsigned int cnt = *count_in;
    int round = 0;
    for (long long base = (long long)blockIdx.x * kJumpSpan; base < cnt;
         base += (long long)gridDim.x * kJumpSpan, round++) {
        uint32_t idx[kJumpPer], a[kJumpPer];
#pragma unroll
        for (int c = 0; c < kJumpPer; c++) {
            const long long j = base + c * kJumpThreads + threadIdx.x;
            idx[c] = j < cnt ? __ldcg(list_in + j) : 0xffffffffu;
        }
#pragma unroll
        for (int c = 0; c < kJumpPer; c++) a[c] = idx[c] != 0xffffffffu ? __ldcg(S + idx[c]) : kFinal;
#pragma unroll 2
        for (int h = 0; h < kHops; h++) {
#pragma unroll
            for (int c = 0; c < kJumpPer; c++)
                if ((a[c] & kFinal) == 0u) a[c] = __ldcg(S + a[c]);
        }
        bool open[kJumpPer];
#pragma unroll
        for (int c = 0; c < kJumpPer; c++) {
            if (idx[c] != 0xffffffffu) __stcg(S + idx[c], a[c]);
            open[c] = (a[c] & kFinal) == 0u;
        }
        block_list_append(open, idx, list_out, count_out, sh, round);
    }
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
                                     void *jump_scratch, long long piece, uint8_t *d_out,
                                     cudaStream_t st)
{
    if (out_hi <= out_lo) return cudaSuccess;
    const DecodeTables t = decode_tables(scratch, n_tokens, P);
    uint32_t *S = reinterpret_cast<uint32_t *>(jump_scratch);
    char *js = reinterpret_cast<char *>(jump_scratch);
    const size_t arr = jump_array_bytes(piece);
    uint32_t *lists[2] = {reinterpret_cast<uint32_t *>(js + arr),
                          reinterpret_cast<uint32_t *>(js + 2 * arr)};
    // open[0]: the build wrote a pointer; open[p + 1]: length of the list pass p wrote
    unsigned int *open = reinterpret_cast<unsigned int *>(js + 3 * arr);
    const long long n_words = (n_in_bytes + 3) / 4;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 8;

    for (long long lo = out_lo; lo < out_hi; lo += piece) {
        const long long hi = lo + piece < out_hi ? lo + piece : out_hi;
        const bool piece_to_end = to_end && hi == out_hi;
        const int n = (int)(hi - lo);
        // depth <= n; a pass with h hops divides it by h + 1 and one more pass finalises.
        // The sweep takes kJumpHops, the list passes kJumpLateHops
        int passes = 0;
        long long depth = n;
        while (depth > 1) {
            const int h = passes == 0 ? kJumpHops : kJumpLateHops;
            depth = (depth + h) / (h + 1);
            passes++;
        }
        passes += 2;
        if (passes > kMaxJumpPasses) passes = kMaxJumpPasses;
        cudaError_t rc = cudaMemsetAsync(open, 0, (kMaxJumpPasses + 2) * sizeof(unsigned int), st);
        if (rc != cudaSuccess) return rc;
        const size_t heads_smem = (size_t)(kJumpThreads / 32) * (1u << P.lb) * 4;
        lz77_jump_build_kernel<<<grid, kJumpThreads, heads_smem, st>>>(
            d_in_words, n_words, n_tokens, P, P.tile_shift, t.tile_tok, t.tile_pos, t.group_pos, lo,
            hi, piece_to_end, d_out, S, open, t.info);
        lz77_jump_sweep_kernel<kJumpHops><<<grid, kJumpThreads, 0, st>>>(S, n, open, lists[0],
                                                                         open + 1);
        for (int p = 1; p < passes; p++)
            lz77_jump_list_kernel<kJumpLateHops><<<grid, kJumpThreads, 0, st>>>(
                S, lists[(p - 1) & 1], open + p, lists[p & 1], open + p + 1);
        lz77_jump_extract_kernel<<<grid, kJumpThreads, 0, st>>>(S, n, d_out + lo, open + passes,
                                                                t.info);
    }
    return cudaGetLastError();
}

}  // namespace lz77
