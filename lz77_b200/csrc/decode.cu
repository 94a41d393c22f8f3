// decode.cu -- token-parallel LZ77 decoder kernels for sm_100a.
//
// Replaces lz77.c:148-197 (decode loop), lz77.c:260-283 (readcode) and
// bitio.c:256-298 (bit-at-a-time reader).  Tokens are fixed width, so token k
// sits at bit 32 + k*T and output offsets are a prefix sum of len+1 -- there is
// no serial bitstream parse.
//
//   pass 1  lz77_decode_scan_kernel   one pass over the tokens (decoupled
//           look-back prefix sum): decoded size, for every output tile the
//           token that contains its first byte, and the output position of
//           every group of 32 tokens.
//   pass 2  lz77_decode_tile_kernel   one CTA per output tile: the tile is
//           assembled in shared memory with one lane per token (literal store +
//           ascending match copy, lz77.c:178-194) and written to HBM with one
//           TMA bulk store.  A match is copied as soon as a per-byte "ready"
//           bitmap shows its source bytes final (dataflow execution: no commit
//           order, no barrier between tokens); sources in earlier tiles
//           (streams the reference encoder wrote reach back SB bytes from
//           anywhere) are read from HBM once that tile has been published.
//           Streams of the block-parallel encoder never leave their tile
//           (tile == encoder block), so every tile decodes independently.
#include "kernels.cuh"

namespace lz77 {

// ---------------------------------------------------------------------------
// pass 1: token length scan
// ---------------------------------------------------------------------------

constexpr int kDsThreads = 256;
constexpr int kDsRows = 16;                          // tokens per lane
constexpr int kDsChunk = kDsThreads * kDsRows;       // tokens per CTA
constexpr unsigned long long kFlagAgg = 1ull << 62;  // chunk aggregate published
constexpr unsigned long long kFlagInc = 2ull << 62;  // inclusive prefix published
constexpr unsigned long long kValMask = (1ull << 62) - 1;

struct DecodeScratch {
    unsigned long long *status;  // look-back state per scan chunk
    long long *tile_tok;         // token containing the first byte of tile j
    long long *tile_pos;         // output position of that token
    uint32_t *group_pos;         // low 32 bits of the output position of token 32g
    unsigned int *tile_done;     // tile j has been written to HBM
    unsigned int *tickets;       // [0] scan chunk ticket, [1] tile ticket
    DecodeInfo *info;
    size_t zero_bytes;           // leading part that must be zeroed per call
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__global__ void __launch_bounds__(kDsThreads)
lz77_decode_scan_kernel(const uint32_t *__restrict__ words, long long n_words, long long n_tokens,
                        Params P, int tile_shift, unsigned long long *status,
                        long long *__restrict__ tile_tok, long long *__restrict__ tile_pos,
                        uint32_t *__restrict__ group_pos, unsigned int *tickets, DecodeInfo *info)
{
    __shared__ long long s_chunk;
    __shared__ unsigned long long s_warp_tot[kDsThreads / 32];
    __shared__ unsigned long long s_prefix;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_chunk = atomicAdd(&tickets[0], 1u);  // chunks start in order
    __syncthreads();
    const long long c = s_chunk;
    const long long warp_base = c * kDsChunk + (long long)warp * (32 * kDsRows);
    const uint32_t len_mask = (1u << P.lb) - 1u;

    // per lane: inclusive scan of len+1 over this warp's 512 tokens
    uint32_t incl[kDsRows];
    uint32_t TOK[kDsRows];  // the token itself (len and off are re-extracted below)
    uint32_t carry = 0;
#pragma unroll
    for (int r = 0; r < kDsRows; r++) {
        const long long k = warp_base + r * 32 + lane;
        uint32_t l1 = 0, tok = 0;
        if (k < n_tokens) {
            tok = load_bits32(words, n_words, kHeaderBits + k * P.tbits);
            l1 = ((tok >> P.ob) & len_mask) + 1u;
        }
        TOK[r] = tok;
        uint32_t x = l1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += t;
        }
        incl[r] = carry + x;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) s_warp_tot[warp] = carry;
    __syncthreads();

    unsigned long long warp_off = 0, chunk_total = 0;
#pragma unroll
    for (int w = 0; w < kDsThreads / 32; w++) {
        const unsigned long long t = s_warp_tot[w];
        if (w < warp) warp_off += t;
        chunk_total += t;
    }

    // decoupled look-back: warp 0 resolves this chunk's exclusive prefix
    if (warp == 0) {
        unsigned long long exclusive = 0;
        if (c > 0) {
            if (lane == 0) atomicExch(&status[c], kFlagAgg | chunk_total);
            long long idx = c - 1 - lane;
            while (true) {
                unsigned long long s = idx >= 0 ? ld_volatile_u64(&status[idx]) : kFlagInc;
                while (__any_sync(0xffffffffu, (s >> 62) == 0)) {
                    if ((s >> 62) == 0) s = ld_volatile_u64(&status[idx]);
                }
                const unsigned inc_mask = __ballot_sync(0xffffffffu, (s >> 62) == 2);
                const unsigned long long val = s & kValMask;
                if (inc_mask) {
                    const int first = __ffs(inc_mask) - 1;
                    exclusive += warp_sum_u64(lane <= first ? val : 0ull);
                    break;
                }
                exclusive += warp_sum_u64(val);
                idx -= 32;
            }
        }
        if (lane == 0) {
            atomicExch(&status[c], kFlagInc | (exclusive + chunk_total));
            s_prefix = exclusive;
        }
    }
    __syncthreads();
    const unsigned long long base_pos = s_prefix + warp_off;

    // tile table: the token that covers byte j << tile_shift
    const long long tile_bytes = 1LL << tile_shift;
#pragma unroll
    for (int r = 0; r < kDsRows; r++) {
        const long long k = warp_base + r * 32 + lane;
        if (k >= n_tokens) continue;
        const uint32_t len = (TOK[r] >> P.ob) & len_mask;
        const uint32_t Lr = len + 1u;
        const long long pos = (long long)(base_pos + incl[r] - Lr);
        // low 32 bits of the output position of tokens 32g .. 32g+31 (the tile
        // kernel only needs positions relative to its tile)
        if (lane == 0) group_pos[k >> 5] = (uint32_t)pos;
        // a source before the start of the token's block: not a stream of the
        // block-parallel encoder (then the tiles of a block pair must run in order)
        if (len > 0 && (long long)(TOK[r] & ((1u << P.ob) - 1u)) > (pos & (P.block - 1)))
            info->cross_block = 1u;
        const long long j = (pos + tile_bytes - 1) >> tile_shift;
        if ((j << tile_shift) < pos + (long long)Lr) {
            tile_tok[j] = k;
            tile_pos[j] = pos;
        }
        if (k == n_tokens - 1) {
            info->n_out = (unsigned long long)(pos + Lr);
        }
    }
}

// The same pass for byte-aligned tokens (T = 24: the default parameters; T = 32: -s 65535
// -l 255), an order of magnitude fewer instructions: a thread owns 16 CONSECUTIVE tokens, so
// the running position inside its run is 16 additions in registers and the CTA needs one
// warp scan instead of one per row of 32 tokens.  The CTA's 4096 tokens are fetched with
// coalesced 32-bit loads (the stream is word aligned at every 4096th token) into shared
// memory, from where each thread reads its 48 or 64 bytes with 128-bit loads (thread
// stride 48 bytes, or 80 with padding: conflict-free per quarter warp).  The per-token
// work left -- tile table, block-containment check -- only runs in the few threads whose
// run touches a tile boundary or the first `window` bytes of a block.
#ifndef LZ77_DS_MINBLOCKS
#define LZ77_DS_MINBLOCKS 8
#endif
template <int kT>
__global__ void __launch_bounds__(kDsThreads, LZ77_DS_MINBLOCKS)
lz77_decode_scan_fast_kernel(const uint32_t *__restrict__ words, long long n_words,
                             long long n_tokens, Params P, int tile_shift,
                             unsigned long long *status, long long *__restrict__ tile_tok,
                             long long *__restrict__ tile_pos, uint32_t *__restrict__ group_pos,
                             unsigned int *tickets, DecodeInfo *info)
{
    constexpr int kPer = kDsRows;                    // consecutive tokens per thread
    constexpr int kTw = kPer * kT / 32;              // words per thread: 12 or 16
    constexpr int kStride = kT == 32 ? 20 : 12;      // words between two threads' runs
    __shared__ __align__(16) uint32_t s_tok[kDsThreads * kStride];
    __shared__ long long s_chunk;
    __shared__ unsigned long long s_warp_tot[kDsThreads / 32];
    __shared__ unsigned long long s_prefix;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_chunk = atomicAdd(&tickets[0], 1u);  // chunks start in order
    __syncthreads();
    const long long c = s_chunk;
    const long long tok0 = c * kDsChunk;
    const long long w0 = 1 + tok0 * kT / 32;  // (4096 tokens are a whole number of words)
    for (int i = threadIdx.x; i < kDsThreads * kTw; i += kDsThreads) {
        const long long w = w0 + i;
        const uint32_t v = w < n_words ? __ldg(words + w) : 0u;
        s_tok[(i / kTw) * kStride + (i % kTw)] = v;
    }
    __syncthreads();

    uint32_t wv[kTw + 1];
    {
        const uint4 *p = reinterpret_cast<const uint4 *>(s_tok + threadIdx.x * kStride);
#pragma unroll
        for (int q = 0; q < kTw / 4; q++) {
            const uint4 v = p[q];
            wv[4 * q] = v.x, wv[4 * q + 1] = v.y, wv[4 * q + 2] = v.z, wv[4 * q + 3] = v.w;
        }
        wv[kTw] = 0u;
    }
    const uint32_t len_mask = (1u << P.lb) - 1u, off_mask = (1u << P.ob) - 1u;
    const long long k_first = tok0 + (long long)threadIdx.x * kPer;
    // token i of the run, from the words in registers (recomputed where it is needed again:
    // keeping all 16 alive costs the registers of two more resident CTAs per SM)
    auto token = [&](int i) -> uint32_t {
        if (kT == 32) return wv[i];
        const int b = 3 * i;
        return __funnelshift_r(wv[b >> 2], wv[(b >> 2) + 1], (b & 3) * 8) & 0xffffffu;
    };
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < kPer; i++)
        if (k_first + i < n_tokens) sum += ((token(i) >> P.ob) & len_mask) + 1u;
    // exclusive prefix of the thread sums inside the CTA
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp_tot[warp] = inc;
    __syncthreads();
    unsigned long long warp_off = 0, chunk_total = 0;
#pragma unroll
    for (int w = 0; w < kDsThreads / 32; w++) {
        const unsigned long long t = s_warp_tot[w];
        if (w < warp) warp_off += t;
        chunk_total += t;
    }
    // decoupled look-back: warp 0 resolves this chunk's exclusive prefix
    if (warp == 0) {
        unsigned long long exclusive = 0;
        if (c > 0) {
            if (lane == 0) atomicExch(&status[c], kFlagAgg | chunk_total);
            long long idx = c - 1 - lane;
            while (true) {
                unsigned long long sv = idx >= 0 ? ld_volatile_u64(&status[idx]) : kFlagInc;
                while (__any_sync(0xffffffffu, (sv >> 62) == 0)) {
                    if ((sv >> 62) == 0) sv = ld_volatile_u64(&status[idx]);
                }
                const unsigned inc_mask = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
                const unsigned long long val = sv & kValMask;
                if (inc_mask) {
                    const int first = __ffs(inc_mask) - 1;
                    exclusive += warp_sum_u64(lane <= first ? val : 0ull);
                    break;
                }
                exclusive += warp_sum_u64(val);
                idx -= 32;
            }
        }
        if (lane == 0) {
            atomicExch(&status[c], kFlagInc | (exclusive + chunk_total));
            s_prefix = exclusive;
        }
    }
    __syncthreads();
    if (k_first >= n_tokens) return;
    const long long p_first = (long long)(s_prefix + warp_off + (unsigned long long)(inc - sum));
    const long long p_end = p_first + (long long)sum;
    if ((threadIdx.x & 1) == 0) group_pos[k_first >> 5] = (uint32_t)p_first;
    if (k_first + kPer >= n_tokens) info->n_out = (unsigned long long)p_end;  // holds the last token

    const long long tile_bytes = 1LL << tile_shift;
    const bool tile_edge = (((p_first + tile_bytes - 1) >> tile_shift) << tile_shift) < p_end;
    const long long q0 = p_first & (P.block - 1);
    const bool near_block_start = q0 < (long long)P.window || q0 + (long long)sum > P.block;
    if (!tile_edge && !near_block_start) return;
    long long pos = p_first;
#pragma unroll
    for (int i = 0; i < kPer; i++) {
        if (k_first + i >= n_tokens) break;
        const uint32_t tk = token(i);
        const uint32_t len = (tk >> P.ob) & len_mask;
        const long long Lr = (long long)len + 1;
        if (len > 0 && (long long)(tk & off_mask) > (pos & (P.block - 1))) info->cross_block = 1u;
        const long long j = (pos + tile_bytes - 1) >> tile_shift;
        if ((j << tile_shift) < pos + Lr) {
            tile_tok[j] = k_first + i;
            tile_pos[j] = pos;
        }
        pos += Lr;
    }
}

// The chunked host path reads the scan's progress without a D2H memcpy: the output
// position behind the last scanned token, bit 63 = a match left its block.
__global__ void lz77_decode_publish_kernel(const DecodeInfo *info,
                                           unsigned long long *host_slot /* mapped pinned */)
{
    *reinterpret_cast<volatile unsigned long long *>(host_slot) =
        info->n_out | (info->cross_block ? 1ull << 63 : 0ull);
    __threadfence_system();
}

// ---------------------------------------------------------------------------
// token-array helpers for sharding one stream across GPUs: tokens
// are fixed width, so a stream splits at any token without parsing
// ---------------------------------------------------------------------------

// out = the header of `words` + tokens [tok_lo, tok_hi) moved down to bit 32
__global__ void lz77_slice_tokens_kernel(const uint32_t *__restrict__ words, long long n_words,
                                         long long bit_lo, long long n_bits,
                                         uint32_t *__restrict__ out, long long out_words)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= out_words) return;
    if (w == 0) {
        out[0] = words[0];
        return;
    }
    const long long b = (w - 1) * 32;  // first payload bit of this word
    uint32_t v = b < n_bits ? load_bits32(words, n_words, bit_lo + b) : 0u;
    if (n_bits - b < 32 && b < n_bits) v &= (1u << (int)(n_bits - b)) - 1u;  // zero padding
    out[w] = v;
}

// The token that holds output byte `pos` (or starts there) and its output position:
// the tile table gives the token covering the tile's first byte, one warp walks on.
__global__ void lz77_token_at_kernel(const uint32_t *__restrict__ words, long long n_words,
                                     long long n_tokens, Params P, int tile_shift,
                                     const long long *__restrict__ tile_tok,
                                     const long long *__restrict__ tile_pos, long long pos,
                                     long long *result /* [2] */)
{
    const int lane = threadIdx.x;
    const uint32_t len_mask = (1u << P.lb) - 1u;
    const long long j = pos >> tile_shift;
    long long k = tile_tok[j], p = tile_pos[j];
    while (true) {
        const long long kk = k + lane;
        unsigned l1 = 0;
        if (kk < n_tokens)
            l1 = ((load_bits32(words, n_words, kHeaderBits + kk * P.tbits) >> P.ob) & len_mask) + 1u;
        unsigned inc = l1;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, inc, dlt);
            if (lane >= dlt) inc += t;
        }
        const unsigned m = __ballot_sync(0xffffffffu, kk < n_tokens && p + (long long)inc > pos);
        if (m) {
            const int first = __ffs(m) - 1;
            if (lane == first) {
                result[0] = kk;
                result[1] = p + inc - l1;
            }
            return;
        }
        p += __shfl_sync(0xffffffffu, inc, 31);
        k += 32;
        if (k >= n_tokens) {  // pos == decoded size
            if (lane == 0) {
                result[0] = n_tokens;
                result[1] = p;
            }
            return;
        }
    }
}

// ---------------------------------------------------------------------------
// pass 2: tile decode
// ---------------------------------------------------------------------------

#ifndef LZ77_DEC_PREFETCH
#define LZ77_DEC_PREFETCH 0
#endif
#ifndef LZ77_DEC_SPINS
#define LZ77_DEC_SPINS 2
#endif
#ifndef LZ77_DEC_SLEEP_NS
#define LZ77_DEC_SLEEP_NS 40
#endif
#ifndef LZ77_DEC_THREADS
#define LZ77_DEC_THREADS 512
#endif
#ifndef LZ77_DEC_WORDCOPY
#define LZ77_DEC_WORDCOPY 1  // matches of <= 16 bytes that do not overlap: word loads, byte stores
#endif
#ifndef LZ77_DEC_EAGER
#define LZ77_DEC_EAGER 1  // copy whatever is ready at once (0: wait for the ready set to settle;
                          // measured 1.05 -> 0.91 ms on 256 MiB of text)
#endif
constexpr int kDecThreadsSmall = LZ77_DEC_THREADS;  // CTA size for 64 KiB tiles (3 CTAs/SM)
constexpr int kDecSpins = LZ77_DEC_SPINS;        // polls without progress before backing off
constexpr unsigned kDecSleepNs = LZ77_DEC_SLEEP_NS;

// ready bitmap: bit i of the tile is set once byte i holds its final value
__device__ __forceinline__ void ready_mark(uint32_t *bits, int a, int b)  // [a, b), a < b
{
    const int w0 = a >> 5, w1 = (b - 1) >> 5;
    const uint32_t first = 0xffffffffu << (a & 31);
    const uint32_t last = 0xffffffffu >> (31 - ((b - 1) & 31));
    if (w0 == w1) {
        atomicOr(bits + w0, first & last);
    } else {
        atomicOr(bits + w0, first);
        for (int w = w0 + 1; w < w1; w++) atomicOr(bits + w, 0xffffffffu);
        atomicOr(bits + w1, last);
    }
}

// advances a over the bytes of [a, b) that are ready; true when all of them are
__device__ __forceinline__ bool ready_test(const volatile uint32_t *bits, int &a, int b)
{
    while (a < b) {
        const int w = a >> 5;
        const int end = min(b, (w + 1) << 5);
        const uint32_t mask = (0xffffffffu << (a & 31)) & (0xffffffffu >> (31 - ((end - 1) & 31)));
        if ((bits[w] & mask) != mask) return false;
        a = end;
    }
    return true;
}

template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
lz77_decode_tile_kernel(const uint32_t *__restrict__ words, long long n_words, long long n_tokens,
                        Params P, int tile_shift, const long long *__restrict__ tile_tok,
                        const long long *__restrict__ tile_pos,
                        const uint32_t *__restrict__ group_pos, long long tile_begin,
                        long long tile_end, long long n_tiles, long long n_out, uint8_t *out,
                        unsigned int *tile_done, unsigned int *ticket, DecodeInfo *info,
                        int pair_mode)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int kWarps = kThreads / 32;
    const int tile_bytes = 1 << tile_shift;
    uint8_t *tile = smem;
    // one bit per tile byte: the byte holds its final value
    uint32_t *ready_bits = reinterpret_cast<uint32_t *>(smem + tile_bytes);

    __shared__ long long s_tile;
    __shared__ int s_next_group;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t off_mask = (1u << P.ob) - 1u;
    const uint32_t len_mask = (1u << P.lb) - 1u;
    const int lit_shift = P.ob + P.lb;

    while (true) {
        if (threadIdx.x == 0) {
            long long t = atomicAdd(ticket, 1u);  // tiles start in ticket order
            if (pair_mode) {
                // a large block is 2^pair_mode tiles and a tile may copy from the tiles of
                // its block before it: hand out the first tile of every block, then the
                // second of every block, ... so no CTA ever waits for a tile of its block
                const long long n = tile_end - tile_begin;
                const int tpb = 1 << pair_mode;
                long long tile = n;
                for (int p = 0; p < tpb; p++) {
                    const long long cnt = n > p ? (n - p + tpb - 1) >> pair_mode : 0;
                    if (t < cnt) {
                        tile = p + (t << pair_mode);
                        break;
                    }
                    t -= cnt;
                }
                t = tile;
            }
            s_tile = tile_begin + t;
            s_next_group = 0;
        }
        __syncthreads();
        const long long j = s_tile;
        if (j >= tile_end) break;

        const long long tile_lo = j << tile_shift;
        long long tile_hi = tile_lo + tile_bytes;
        if (tile_hi > n_out) tile_hi = n_out;
        const int tile_len = (int)(tile_hi - tile_lo);
        const long long k0 = tile_tok[j];
        long long k_end = n_tokens;
        if (j + 1 < n_tiles) k_end = tile_tok[j + 1] + (tile_pos[j + 1] < tile_hi ? 1 : 0);
        // groups of 32 tokens are global (token 32g .. 32g+31); pass 1 left the output
        // position of each in group_pos.  The tile's first and last group may hold
        // tokens of the neighbouring tiles: those lanes only feed the prefix sum.
        const long long g_first = k0 >> 5;
        const int n_groups = (int)(((k_end + 31) >> 5) - g_first);
        const uint32_t tile_lo32 = (uint32_t)tile_lo;
        // token addressing relative to the tile's first group: 32-bit arithmetic per token
        // (a tile holds at most 2^17 tokens of at most 32 bits)
        const long long bit0 = kHeaderBits + (g_first << 5) * P.tbits;
        const uint32_t *tile_words = words + (bit0 >> 5);
        const int bit0_low = (int)(bit0 & 31);
        const long long words_left = n_words - (bit0 >> 5);
        const int n_words_rel = words_left > 0x7fffffff ? 0x7fffffff : (int)words_left;
        const int k0_rel = (int)(k0 - (g_first << 5)), k_end_rel = (int)(k_end - (g_first << 5));

        for (int i = threadIdx.x; i < (tile_bytes >> 5); i += kThreads) ready_bits[i] = 0u;
        __syncthreads();

        // ---- one lane per token, groups handed out in order -----------------
        while (true) {
            int gi = 0;
            if (lane == 0) gi = atomicAdd(&s_next_group, 1);
            gi = __shfl_sync(0xffffffffu, gi, 0);
            if (gi >= n_groups) break;

            const int k = (gi << 5) + lane;  // relative to the tile's first group
            const bool valid = k >= k0_rel && k < k_end_rel;
            uint32_t tok = 0;
            if (k < k_end_rel) {
                const int b = bit0_low + k * P.tbits;
                const int w = b >> 5, sft = b & 31;
                const uint32_t lo = w < n_words_rel ? __ldg(tile_words + w) : 0u;
                const uint32_t hi = (sft != 0 && w + 1 < n_words_rel) ? __ldg(tile_words + w + 1) : 0u;
                tok = __funnelshift_r(lo, hi, sft);
            }
#if LZ77_DEC_PREFETCH
            {
                // the tokens this warp will most likely be handed next (one group per warp
                // and round): pull their line towards the SM while this group is decoded
                const int kp = k + (kWarps << 5);
                if (lane == 0 && kp < k_end_rel)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(
                        tile_words + ((bit0_low + kp * P.tbits) >> 5)));
            }
#endif
            const int off = (int)(tok & off_mask);
            const int len = (int)((tok >> P.ob) & len_mask);
            const uint32_t lit = (tok >> lit_shift) & 0xffu;
            const int l1 = k < k_end_rel ? len + 1 : 0;  // lanes before k0 still count in the sum
            int inc = l1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            const int pos_rel = (int)(__ldg(group_pos + g_first + gi) - tile_lo32) + inc - l1;
            // the part of this token that lies in the tile
            const int d_lo = max(pos_rel, 0), d_hi = min(pos_rel + l1, tile_len);

            if (valid) {  // literal, lz77.c:189-194
                const int d = pos_rel + len;
                if (d >= 0 && d < tile_len) tile[d] = (uint8_t)lit;
            }
            bool pending = valid && len > 0;
            if (pending && (off == 0 || (long long)off > tile_lo + pos_rel)) {
                info->error = 1;  // source before the start of the output
                pending = false;
            }
            // every source byte lies in [s_rel, e_rel), strictly before the token
            // (a self-overlapping match repeats its first `off` bytes)
            const int s_rel = pos_rel - off;
            const int e_rel = s_rel + min(len, off);
            int chk = max(s_rel, 0);          // in-tile source bytes below chk are known ready
            bool ext = pending && s_rel < 0;  // reaches into earlier tiles
            unsigned int ta = 0, tb = 0;
            if (ext) {
                ta = (unsigned int)((tile_lo + s_rel) >> tile_shift);
                tb = (unsigned int)((tile_lo + min(e_rel, 0) - 1) >> tile_shift);
            }
            __threadfence_block();
            if (valid && !pending && d_lo < d_hi) ready_mark(ready_bits, d_lo, d_hi);
            __syncwarp();

            // token and sources entirely inside the tile: the common, cheap copy
            const bool fast = pos_rel >= 0 && s_rel >= 0 && pos_rel + len <= tile_len;
            bool ready = false;  // sources final, copy still to do
            // in-tile sources of at most 32 bytes: one 64-bit window of the ready bitmap
            const int nsrc = e_rel - chk;
            const bool fastpoll = nsrc <= 32;
            const int pw = chk >> 5, psh = chk & 31;
            const uint32_t pmask = nsrc >= 32 ? 0xffffffffu : nsrc <= 0 ? 0u : (1u << nsrc) - 1u;
            const volatile uint32_t *vbits = ready_bits;
            unsigned prev_rm = 0;
            int spins = 0;
            while (true) {
                const unsigned um = __ballot_sync(0xffffffffu, pending);
                if (!um) break;
                if (pending && !ready) {  // cheap poll of the ready bitmap
                    if (fastpoll) {
                        const uint32_t lo = vbits[pw], hi = vbits[pw + 1];
                        ready = (__funnelshift_r(lo, hi, psh) & pmask) == pmask;
                    } else {
                        ready = ready_test(ready_bits, chk, e_rel);
                    }
                    if (ready && ext) {
                        ready = ld_acquire_u32(&tile_done[ta]) != 0u &&
                                ld_acquire_u32(&tile_done[tb]) != 0u;
                        if (ready) ext = false;
                    }
                }
                const unsigned rm = __ballot_sync(0xffffffffu, ready);
                // copy in as few divergent passes as possible: when every pending
                // lane is ready, or when the ready set stopped growing
                const bool go = rm != 0u && (LZ77_DEC_EAGER || rm == um || rm == prev_rm);
                prev_rm = rm;
                if (go) {
                    if (ready) {  // ascending byte copy, lz77.c:178-188
                        __threadfence_block();  // acquire: bitmap read -> source bytes
                        if (fast) {
                            const volatile uint8_t *src = tile + s_rel;
                            uint8_t *dst = tile + pos_rel;
                            if (LZ77_DEC_WORDCOPY && off >= len && len <= 16) {
                                // source and destination do not overlap, at most 16 bytes:
                                // five aligned words up front (one round of loads instead of
                                // a load per byte), then predicated byte stores
                                // (four bytes at a time, the next aligned word loaded only when
                                // the match goes on: two thirds of the matches end within four
                                // bytes, and sixteen predicated byte stores cost every pass ~80
                                // instructions whatever the lengths)
                                const volatile uint32_t *sw =
                                    reinterpret_cast<const volatile uint32_t *>(tile + (s_rel & ~3));
                                const int sh = s_rel * 8;  // (the funnel shift wraps: bits 3..4 count)
                                uint32_t lo = sw[0], hi = sw[1];
                                uint32_t v = __funnelshift_r(lo, hi, sh);
                                dst[0] = (uint8_t)v;  // (len >= 1 here)
                                if (len > 1) dst[1] = (uint8_t)(v >> 8);
                                if (len > 2) dst[2] = (uint8_t)(v >> 16);
                                if (len > 3) dst[3] = (uint8_t)(v >> 24);
#pragma unroll
                                for (int c = 1; c < 4; c++) {
                                    if (len <= 4 * c) break;
                                    lo = hi, hi = sw[c + 1];
                                    v = __funnelshift_r(lo, hi, sh);
                                    uint8_t *d = dst + 4 * c;
                                    d[0] = (uint8_t)v;
                                    if (len > 4 * c + 1) d[1] = (uint8_t)(v >> 8);
                                    if (len > 4 * c + 2) d[2] = (uint8_t)(v >> 16);
                                    if (len > 4 * c + 3) d[3] = (uint8_t)(v >> 24);
                                }
                            } else if (off >= len) {
                                for (int i = 0; i < len; i++) dst[i] = src[i];
                            } else {
                                int r = 0;
                                for (int i = 0; i < len; i++) {
                                    dst[i] = src[r];
                                    if (++r == off) r = 0;
                                }
                            }
                        } else {
                            int r = 0;
                            for (int i = 0; i < len; i++) {
                                const int a = s_rel + r;
                                const uint8_t c = a >= 0 ? const_cast<volatile uint8_t *>(tile)[a]
                                                         : __ldcg(out + (tile_lo + a));
                                const int d = pos_rel + i;
                                if (d >= 0 && d < tile_len) tile[d] = c;
                                if (++r == off) r = 0;
                            }
                        }
                        __threadfence_block();
                        if (d_lo < d_hi) ready_mark(ready_bits, d_lo, d_hi);
                        pending = false;
                        ready = false;
                    }
                    __syncwarp();
                    prev_rm = 0;
                    spins = 0;
                } else if (rm == 0u && ++spins > kDecSpins) {
                    __nanosleep(kDecSleepNs);  // leave the issue slots to the warps we wait for
                }
            }
        }
        __syncthreads();

        // ---- flush the tile: one TMA bulk store (shared -> global) -----------
        {
            uint8_t *dst = out + tile_lo;
            const int n16 = tile_len & ~15;
            fence_proxy_async();  // the tile was written through the generic proxy
            __syncthreads();
            if (threadIdx.x == 0 && n16 > 0) {
                tma_store_1d(dst, tile, (uint32_t)n16);
                tma_store_commit_wait();  // complete (and visible) before the flag below
            }
            for (int i = n16 + threadIdx.x; i < tile_len; i += kThreads) dst[i] = tile[i];
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) st_release_u32(&tile_done[j], 1u);
    }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------

int decode_tile_bytes(const Params &P)
{
    return 1 << P.tile_shift;  // the encoder block, or a quarter of a 512 KiB block
}

static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

static DecodeScratch carve_decode(void *scratch, long long n_tokens, const Params &P)
{
    const int tile_shift = P.tile_shift;
    const long long n_chunks = (n_tokens + kDsChunk - 1) / kDsChunk;
    const long long max_out = n_tokens << P.lb;  // len + 1 <= 2^lb
    const long long max_tiles = (max_out >> tile_shift) + 2;
    char *p = (char *)scratch;
    DecodeScratch s;
    s.info = (DecodeInfo *)p;
    p += 256;
    s.tickets = (unsigned int *)p;
    p += 256;
    s.status = (unsigned long long *)p;
    p += al256((size_t)n_chunks * 8);
    s.tile_done = (unsigned int *)p;
    p += al256((size_t)max_tiles * 4);
    s.zero_bytes = (size_t)(p - (char *)scratch);
    s.tile_tok = (long long *)p;
    p += al256((size_t)max_tiles * 8);
    s.tile_pos = (long long *)p;
    p += al256((size_t)max_tiles * 8);
    s.group_pos = (uint32_t *)p;
    p += al256((size_t)((n_tokens + 31) / 32) * 4);
    return s;
}

DecodeTables decode_tables(void *scratch, long long n_tokens, const Params &P)
{
    const DecodeScratch s = carve_decode(scratch, n_tokens, P);
    return DecodeTables{s.tile_tok, s.tile_pos, s.group_pos, s.info};
}

size_t decode_scratch_bytes(long long n_tokens, const Params &P)
{
    const long long n_chunks = (n_tokens + kDsChunk - 1) / kDsChunk;
    const long long max_tiles = ((n_tokens << P.lb) >> P.tile_shift) + 2;
    return 512 + al256((size_t)n_chunks * 8) + al256((size_t)max_tiles * 4) +
           2 * al256((size_t)max_tiles * 8) + al256((size_t)((n_tokens + 31) / 32) * 4) + 1024;
}

int decode_launch_count(bool with_copy) { return with_copy ? 2 : 1; }

// Pass 1 over tokens [tok_begin, tok_end) of a stream of n_tokens tokens, of which
// the first n_in_bytes bytes are resident.  tok_begin must be a multiple of the
// scan chunk; the first call (tok_begin == 0) zeroes the state.  After the call
// info->n_out holds the output position behind token tok_end - 1.
cudaError_t launch_decode_scan_range(const uint32_t *d_in_words, long long n_in_bytes,
                                     long long n_tokens, long long tok_begin, long long tok_end,
                                     const Params &P, void *scratch, DecodeInfo **d_info,
                                     cudaStream_t st, unsigned long long *host_n_out)
{
    DecodeScratch s = carve_decode(scratch, n_tokens, P);
    *d_info = s.info;
    if (tok_begin == 0) {
        cudaError_t rc = cudaMemsetAsync(scratch, 0, s.zero_bytes, st);
        if (rc != cudaSuccess) return rc;
    }
    const long long n_chunks = (tok_end - tok_begin + kDsChunk - 1) / kDsChunk;
    const long long n_words = (n_in_bytes + 3) / 4;
    if (n_chunks > 0) {
        auto kern = P.tbits == 24   ? lz77_decode_scan_fast_kernel<24>
                    : P.tbits == 32 ? lz77_decode_scan_fast_kernel<32>
                                    : lz77_decode_scan_kernel;  // any other width
        kern<<<(unsigned)n_chunks, kDsThreads, 0, st>>>(d_in_words, n_words, tok_end, P,
                                                        P.tile_shift, s.status, s.tile_tok,
                                                        s.tile_pos, s.group_pos, s.tickets, s.info);
    }
    if (host_n_out) lz77_decode_publish_kernel<<<1, 1, 0, st>>>(s.info, host_n_out);
    return cudaGetLastError();
}

long long decode_scan_granule() { return kDsChunk; }

cudaError_t launch_slice_tokens(const uint32_t *d_in_words, long long n_in_bytes,
                                long long tok_lo, long long tok_hi, const Params &P,
                                uint32_t *d_out_words, long long out_words, cudaStream_t st)
{
    const long long n_words = (n_in_bytes + 3) / 4;
    const int threads = 256;
    const long long blocks = (out_words + threads - 1) / threads;
    lz77_slice_tokens_kernel<<<(unsigned)blocks, threads, 0, st>>>(
        d_in_words, n_words, kHeaderBits + tok_lo * P.tbits, (tok_hi - tok_lo) * P.tbits,
        d_out_words, out_words);
    return cudaGetLastError();
}

// after launch_decode_scan; result = 2 x int64 in device memory
cudaError_t launch_token_at(const uint32_t *d_in_words, long long n_in_bytes, long long n_tokens,
                            long long pos, const Params &P, void *scratch, long long *d_result,
                            cudaStream_t st)
{
    DecodeScratch s = carve_decode(scratch, n_tokens, P);
    lz77_token_at_kernel<<<1, 32, 0, st>>>(d_in_words, (n_in_bytes + 3) / 4, n_tokens, P,
                                           P.tile_shift, s.tile_tok, s.tile_pos, pos, d_result);
    return cudaGetLastError();
}

cudaError_t launch_decode_scan(const uint32_t *d_in_words, long long n_in_bytes,
                               long long n_tokens, const Params &P, void *scratch,
                               DecodeInfo **d_info, cudaStream_t st)
{
    return launch_decode_scan_range(d_in_words, n_in_bytes, n_tokens, 0, n_tokens, P, scratch,
                                    d_info, st, nullptr);
}

// Pass 2 over output tiles [tile_begin, tile_end).  `last` is false for a partial
// run of a chunked decode: then every tile of the range is complete and the
// token that starts tile_end has been scanned; launch_idx (< 60) selects the
// ticket slot of this launch.
cudaError_t launch_decode_tiles_range(const uint32_t *d_in_words, long long n_in_bytes,
                                      long long n_tokens, long long tile_begin,
                                      long long tile_end, bool last, long long n_out,
                                      int launch_idx, int pair_mode, const Params &P,
                                      void *scratch, uint8_t *d_out, cudaStream_t st)
{
    DecodeScratch s = carve_decode(scratch, n_tokens, P);
    const int tile_shift = P.tile_shift;
    const long long tile_bytes = 1LL << tile_shift;
    const long long n_words = (n_in_bytes + 3) / 4;
    const long long n_run = tile_end - tile_begin;
    if (n_run <= 0) return cudaSuccess;
    const long long big = 1LL << 60;
    const long long n_tiles_total = last ? tile_end : big;
    const long long n_out_eff = last ? n_out : big;
    unsigned int *ticket = s.tickets + 1 + launch_idx;
    const size_t smem = (size_t)tile_bytes + (size_t)(tile_bytes >> 3) + 16;  // tile + ready bitmap

    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (tile_bytes <= 65536) {
        auto kern = lz77_decode_tile_kernel<kDecThreadsSmall, 3>;
        cudaError_t rc =
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return rc;
        long long grid = (long long)sms * 3;
        if (grid > n_run) grid = n_run;
        kern<<<(unsigned)grid, kDecThreadsSmall, smem, st>>>(d_in_words, n_words, n_tokens, P, tile_shift,
                                                 s.tile_tok, s.tile_pos, s.group_pos, tile_begin,
                                                 tile_end,
                                                 n_tiles_total, n_out_eff, d_out, s.tile_done,
                                                 ticket, s.info, pair_mode);
    } else {
        auto kern = lz77_decode_tile_kernel<1024, 1>;
        cudaError_t rc =
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return rc;
        long long grid = sms;
        if (grid > n_run) grid = n_run;
        kern<<<(unsigned)grid, 1024, smem, st>>>(d_in_words, n_words, n_tokens, P, tile_shift,
                                                  s.tile_tok, s.tile_pos, s.group_pos, tile_begin,
                                                 tile_end,
                                                  n_tiles_total, n_out_eff, d_out, s.tile_done,
                                                  ticket, s.info, pair_mode);
    }
    return cudaGetLastError();
}

cudaError_t launch_decode_copy(const uint32_t *d_in_words, long long n_in_bytes,
                               long long n_tokens, long long n_out, bool cross_block,
                               const Params &P, void *scratch, void *jump_scratch,
                               long long jump_piece, uint8_t *d_out, cudaStream_t st)
{
    // matches that leave their block (a stream of the reference encoder): the tiles
    // would form a chain, resolve the copies by pointer jumping instead
    if (cross_block)
        return launch_decode_jump_range(d_in_words, n_in_bytes, n_tokens, 0, n_out, true, P,
                                        scratch, jump_scratch, jump_piece, d_out, st);
    // phase order only for streams whose matches stay inside their (multi-tile) block
    const int pair_mode = (P.block_shift > P.tile_shift && !cross_block)
                              ? P.block_shift - P.tile_shift : 0;
    const long long tile_bytes = 1LL << P.tile_shift;
    const long long n_tiles = (n_out + tile_bytes - 1) >> P.tile_shift;
    return launch_decode_tiles_range(d_in_words, n_in_bytes, n_tokens, 0, n_tiles, true, n_out, 0,
                                     pair_mode, P, scratch, d_out, st);
}

}  // namespace lz77
