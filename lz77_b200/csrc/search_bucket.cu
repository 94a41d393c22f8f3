// search_bucket.cu -- longest-match search over position buckets (sm_100a).
//
// Second-generation replacement of the reference's BST match finder
// (tree.c:62-260).  Same result as the exhaustive window scan in encode.cu --
// the longest match, farthest offset among the longest -- at a fraction of the
// instructions: instead of filtering every window byte per token, each CTA
// first groups the positions of its staged tile by a key of their first two
// bytes, then a token only verifies the positions that share its key.
//
//   build (per tile, all warps)
//     key(q) = (x[q] & 31) << 5 | x[q+1] & 31             (1024 buckets)
//     stable counting sort of the tile's positions by key: per-warp histogram
//     of a contiguous chunk (packed 16-bit counters, shared-memory atomics),
//     column scan across warps + bucket scan, then an in-order scatter whose
//     intra-row ranks come from MATCH.ANY -- so every bucket lists its
//     positions in ascending order.
//   search (per token, one warp)
//     bucket of the lookahead's key -> warp-ary lower bound of the window start
//     -> 32 candidates per round, oldest first, verified against the target
//     held in registers -> (length, oldest start) reduced with REDUX; stops at
//     the first maximum-length match.
//     A match of length 1 can sit in many buckets; when nothing longer exists the
//     warp scans the staged window forward (SWAR byte compare, 512 bytes per
//     step) for the oldest occurrence of the first byte.
//   emit (fused path, 24-bit tokens: the default parameters)
//     replaces writecode() lz77.c:246-252 + bitIO_write() bitio.c:203-239 inside the
//     search kernel.  A warp keeps its segment's tokens in shared memory (the counter
//     area of the build, dead by then; tokens beyond 512 per segment -- incompressible
//     data -- spill to a small per-CTA area in global memory).  When the tile is parsed
//     its token count enters a decoupled look-back over the tiles (tiles are handed out
//     by ticket, so every predecessor is running or done), which yields the tile's place
//     in the stream; the CTA packs its tokens to 3 bytes each in shared memory -- at the
//     same offset modulo 16 as their place in the stream -- and writes them with 128-bit
//     stores, single bytes at the two ends (T is a multiple of 8: no tile shares a byte
//     with its neighbour, so nothing has to be zeroed or merged).  No token ever
//     travels through HBM unpacked.  Other token widths keep the generic path: 32-bit
//     tokens to scratch, count scan, lz77_pack_kernel (encode.cu).
#include "kernels.cuh"
#include "match.cuh"

namespace lz77 {

// Bucket key: the low kKeyBits bits of each of the two bytes.  For lowercase text
// this is a perfect hash of the byte pair (one bucket per digram); for binary
// data it spreads the 65536 pairs evenly.
// tuning knobs, measured on 256 MiB of text (B200): rows of the scatter issued in
// batches of 1 / 2 / 4 / 8 -> 9.01 / 9.28 / 9.52 / 9.79 ms (MATCH.ANY back to back stalls
// the pipe); reading bytes 8..15 of the target on demand -> 9.15 -> 9.01 ms
#ifndef LZ77_SCATTER_BATCH
#define LZ77_SCATTER_BATCH 1
#endif
#ifndef LZ77_BALLOT_RANK
#define LZ77_BALLOT_RANK 0
#endif
#ifndef LZ77_LAZY_TGT
#define LZ77_LAZY_TGT 1
#endif
#ifndef LZ77_KEY_BITS
#define LZ77_KEY_BITS 5
#endif
constexpr int kKeyBits = LZ77_KEY_BITS;
constexpr int kBuckets = 1 << (2 * kKeyBits);
#ifndef LZ77_BACKWALK
#define LZ77_BACKWALK 1  // walk a bucket backwards from the position's own slot (0: forwards from
                         // the bucket start / the window's lower bound)
#endif
#ifndef LZ77_LINEAR_SCAN
#define LZ77_LINEAR_SCAN 128
#endif
#ifndef LZ77_BUILD_ONLY
#define LZ77_BUILD_ONLY 0  // (timing experiments) stage + build only, no parse: the output is invalid
#endif
#ifndef LZ77_DEFER_STORE
#define LZ77_DEFER_STORE 1  // (token loop 2) store a token one pass later, behind the next token's loads
#endif
#ifndef LZ77_TOKLOOP
#define LZ77_TOKLOOP 2  // 2: token loop written against the ALU pipe (packed running best, uniform
                        // loop control); 1: the round-1 loop
#endif
constexpr int kLinearScan = LZ77_LINEAR_SCAN;  // buckets up to this size are scanned from their start

__device__ __forceinline__ int pair_key(uint32_t b0, uint32_t b1)
{
    return (int)(((b0 & ((1u << kKeyBits) - 1u)) << kKeyBits) | (b1 & ((1u << kKeyBits) - 1u)));
}

__device__ __forceinline__ int bucket_key(const uint8_t *smem, int i)
{
    return pair_key(smem[i], smem[i + 1]);
}

// ---- shared-memory reads by 32-bit shared address --------------------------
// Inside the token loop every read goes through an explicit shared-window
// address computed once per kernel: with generic pointers the compiler
// re-derives the CTA's shared window base (S2R SR_CgaCtaId + LEA) in front of
// every access sequence.
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// match_len<false>() of match.cuh on shared addresses (sdata = shared address of the
// staged bytes): lookaheads longer than 16 bytes.  The first 16 bytes are compared with
// the registers, the rest word by word.
__device__ __forceinline__ int match_len_long(uint32_t sdata, int q, int p0,
                                              const uint32_t (&tgt)[4], int max_len)
{
    const uint32_t w = sdata + (uint32_t)(q & ~3);
    const int sh = (q & 3) * 8;
    const uint32_t a0 = lds32(w), a1 = lds32(w + 4);
    uint32_t x = __funnelshift_r(a0, a1, sh) ^ tgt[0];
    if (x) return min((__ffs(x) - 1) >> 3, max_len);
    const uint32_t a2 = lds32(w + 8);
    x = __funnelshift_r(a1, a2, sh) ^ tgt[1];
    if (x) return min(4 + ((__ffs(x) - 1) >> 3), max_len);
    const uint32_t a3 = lds32(w + 12);
    x = __funnelshift_r(a2, a3, sh) ^ tgt[2];
    if (x) return min(8 + ((__ffs(x) - 1) >> 3), max_len);
    uint32_t a = lds32(w + 16);
    x = __funnelshift_r(a3, a, sh) ^ tgt[3];
    if (x) return min(12 + ((__ffs(x) - 1) >> 3), max_len);
    int l = 16;
    uint32_t wa = w + 20;
    while (l < max_len) {
        const uint32_t b = lds32(wa);
        wa += 4;
        const int pi = p0 + l;
        const uint32_t pw = sdata + (uint32_t)(pi & ~3);
        const uint32_t t = __funnelshift_r(lds32(pw), lds32(pw + 4), (pi & 3) * 8);
        x = __funnelshift_r(a, b, sh) ^ t;
        if (x) {
            l += (__ffs(x) - 1) >> 3;
            break;
        }
        l += 4;
        a = b;
    }
    return min(l, max_len);
}

// One candidate per lane of a round, LA <= 16.  The first four bytes are compared
// without a branch by every lane -- a lane without a candidate (`in` false) reads the
// target itself and is masked afterwards -- because most candidates end there and a
// divergent round costs every path once; only lanes whose first word matches go on.
__device__ __forceinline__ int round_match_len(uint32_t sdata, int q, bool in, int p0,
                                               uint32_t tgt0, uint32_t tgt1, uint32_t tgt2,
                                               uint32_t tgt3, int max_len)
{
    const int qq = in ? q : p0;
    const uint32_t w = sdata + (uint32_t)(qq & ~3);
    const int sh = (qq & 3) * 8;
    const uint32_t a0 = lds32(w), a1 = lds32(w + 4);
    uint32_t x = __funnelshift_r(a0, a1, sh) ^ tgt0;
    int l = (int)min((uint32_t)(__ffs(x) - 1) >> 3, 4u);  // __ffs(0) - 1 wraps: 4
    if (in && x == 0u) {
        const uint32_t a2 = lds32(w + 8);
        x = __funnelshift_r(a1, a2, sh) ^ tgt1;
        if (x) {
            l = 4 + ((__ffs(x) - 1) >> 3);
        } else {
            // bytes 8..15 of the target are rarely needed: read them here, not per token
            const uint32_t pw = sdata + (uint32_t)(p0 & ~3);
            const int psh = (p0 & 3) * 8;
            const uint32_t t2 = lds32(pw + 8), t3 = lds32(pw + 12), t4 = lds32(pw + 16);
            const uint32_t a3 = lds32(w + 12);
            x = __funnelshift_r(a2, a3, sh) ^ (LZ77_LAZY_TGT ? __funnelshift_r(t2, t3, psh) : tgt2);
            if (x) {
                l = 8 + ((__ffs(x) - 1) >> 3);
            } else {
                const uint32_t a4 = lds32(w + 16);
                x = __funnelshift_r(a3, a4, sh) ^ (LZ77_LAZY_TGT ? __funnelshift_r(t3, t4, psh) : tgt3);
                l = x ? 12 + ((__ffs(x) - 1) >> 3) : 16;
            }
        }
    }
    return in ? min(l, max_len) : 0;
}

// The same compare for the packed-key token loop (LZ77_TOKLOOP 2): no select in front of the
// loads -- a lane without a candidate carries q = 0, reads the first staged bytes and is
// masked by the caller -- and no clamp of the first level (every lane whose first word
// matches is overwritten by the deeper levels).  The result is unspecified when `in` is false.
__device__ __forceinline__ int cand_match_len(uint32_t sdata, int q, bool in, int p0, uint32_t tgt0,
                                              uint32_t tgt1)
{
    const uint32_t w = sdata + (uint32_t)(q & ~3);
    const int sh = q * 8;  // (the funnel shift wraps: only bits 3..4 count)
    // (ballot-predicated loads -- only lanes in front of the first entry that left the window
    // load -- cost fewer bank conflicts but put the ballot in front of the loads: 6.82 vs 6.69 ms)
    const uint32_t a0 = lds32(w), a1 = lds32(w + 4);
    uint32_t x = __funnelshift_r(a0, a1, sh) ^ tgt0;
    int l = (__ffs(x) - 1) >> 3;
    if (in && x == 0u) {
        const uint32_t a2 = lds32(w + 8);
        x = __funnelshift_r(a1, a2, sh) ^ tgt1;
        if (x) {
            l = 4 + ((__ffs(x) - 1) >> 3);
        } else {
            const uint32_t pw = sdata + (uint32_t)(p0 & ~3);
            const int psh = p0 * 8;
            const uint32_t t2 = lds32(pw + 8), t3 = lds32(pw + 12), t4 = lds32(pw + 16);
            const uint32_t a3 = lds32(w + 12);
            x = __funnelshift_r(a2, a3, sh) ^ __funnelshift_r(t2, t3, psh);
            if (x) {
                l = 8 + ((__ffs(x) - 1) >> 3);
            } else {
                const uint32_t a4 = lds32(w + 16);
                x = __funnelshift_r(a3, a4, sh) ^ __funnelshift_r(t3, t4, psh);
                l = x ? 12 + ((__ffs(x) - 1) >> 3) : 16;
            }
        }
    }
    return l;
}

// first index in [0, n) of the ascending uint16 list at shared address se whose
// value is >= lo (n if none); uniform across the kLanes lanes of a group
template <int kLanes>
__device__ __forceinline__ int group_lower_bound_s(uint32_t se, int n, int lo, int sl,
                                                   unsigned gmask, int gshift)
{
    int base = 0, cnt = n;
    while (cnt > kLanes) {
        const int step = (cnt + kLanes - 1) / kLanes;
        const int idx = base + sl * step;
        const bool ge = idx < base + cnt ? (int)lds16(se + 2u * idx) >= lo : true;
        const unsigned m = __ballot_sync(gmask, ge) >> gshift;
        const int first = m ? __ffs(m) - 1 : kLanes;
        if (first == 0) return base;
        const int nb = base + (first - 1) * step + 1;
        const int ne = min(base + cnt, base + first * step + 1);
        base = nb;
        cnt = ne - nb;
    }
    const int idx = base + sl;
    const bool ge = idx < base + cnt ? (int)lds16(se + 2u * idx) >= lo : true;
    const unsigned m = __ballot_sync(gmask, ge) >> gshift;
    return base + (m ? __ffs(m) - 1 : kLanes);
}

// kLanes lanes (a "group": 32, 16 or 8) cooperate on one parse segment, so a warp
// parses 32 / kLanes segments side by side.  Measured on 256 MiB of text (B200), the same
// 8 KiB tile throughout: one warp per 1 KiB segment 8.7 ms; groups of 16 / 8 lanes on 512- /
// 256-byte segments with this loop 8.6 / 9.0 ms (the groups of a warp do not stay in step);
// the same groups driven in LOCKSTEP by a per-group state machine (warp-uniform branches,
// candidates walked backwards from the position's own bucket slot, so no window lower bound)
// 10.5 / 9.6 ms although byte-exact: a pass of the state machine costs the sum of all the
// steps any group is in, and a warp waits for the slowest of its four segments.  One warp
// per segment stays; the fused path is written for kLanes == 32.
#ifndef LZ77_PARSE_MINBLOCKS
#define LZ77_PARSE_MINBLOCKS 1
#endif
#ifndef LZ77_PARSE_WARPS
#define LZ77_PARSE_WARPS 8
#endif
#ifndef LZ77_PARSE_LANES
#define LZ77_PARSE_LANES 32
#endif

constexpr int kTokBuf = 512;                     // tokens per segment kept in shared memory
constexpr int kTokSpill = kSegBytes - kTokBuf;   // the rest of a worst-case segment: global
constexpr int kRing = 3;                         // packed tiles a CTA may have parked
constexpr int kRingSlot = 3 * LZ77_PARSE_WARPS * kSegBytes + 256;  // worst case: a token per byte
constexpr unsigned long long kLbAgg = 1ull << 62, kLbInc = 2ull << 62, kLbVal = (1ull << 62) - 1;

struct FusedEmit {                 // outputs of the fused path (kFused)
    uint8_t *out;                  // the stream (header + tokens), 16-byte aligned
    unsigned long long *status;    // look-back state, one word per tile of the whole call
    unsigned int *ticket;          // tile ticket of this launch
    unsigned long long *total;     // tokens up to and including this launch
    unsigned long long *host_total;  // the same in mapped pinned memory (may be null)
    uint32_t *spill;               // gridDim.x * kWarps * kTokSpill tokens
    uint8_t *ring;                 // gridDim.x * kRing slots of kRingSlot bytes
    long long tile0;               // index of this launch's first tile in the whole call
    int write_header;
};

template <bool kSmallLA, int kWarps, int kLanes, typename PosT, bool kSortedGlobal, bool kFused,
          int kSeg>
__global__ void __launch_bounds__(kWarps * 32, kFused ? 4 : LZ77_PARSE_MINBLOCKS)
lz77_parse_bucket_kernel(const uint8_t *__restrict__ in, long long n, long long pre, Params P,
                         int hist_cap, long long n_tiles, uint32_t *__restrict__ tok_tmp,
                         uint32_t *__restrict__ seg_ntok, PosT *sorted_global,
                         long long sorted_stride, FusedEmit F)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    __shared__ long long s_tile;                 // fused: the tile this CTA drew
    __shared__ int s_cnt[kWarps];                // fused: tokens per segment of the tile
    __shared__ unsigned long long s_excl;        // fused: tokens in front of the tile being resolved
    __shared__ long long s_pend_tile[kRing];     // fused: parked tiles (index in the whole call)
    __shared__ int s_pend_k[kRing];              //        and their token counts
    __shared__ int s_ok;
    static_assert(!kFused || kLanes == 32, "the fused emit is written for one warp per segment");

    constexpr int kThreads = kWarps * 32;
    constexpr int kSegsPerWarp = 32 / kLanes;
    constexpr int kCntWords = kBuckets * (kWarps / 2);  // two 16-bit counters per word
    constexpr int tile_bytes = kWarps * kSegsPerWarp * kSeg;
    const int data_cap = hist_cap + tile_bytes + 64;
    // the backward candidate walk needs the counter area for its slot table: not with the
    // fused emit (its token buffers live there), and only for LA <= 16 (the walk's compare)
    constexpr bool kBackWalk = LZ77_BACKWALK != 0 && kSmallLA && !kFused && kLanes == 32;
    // sentinel layout: [32 zero entries][bucket 0: sentinel 0, entries][bucket 1: ...]; the
    // bucket starts only exist folded into the scatter's counters, there is no bstart array
    constexpr bool kSent = kBackWalk && LZ77_TOKLOOP == 2;  // (the token loop 2 layout)
    PosT *bstart = reinterpret_cast<PosT *>(smem + ((data_cap + 15) & ~15));
    uint32_t *cnt = kSent ? reinterpret_cast<uint32_t *>(bstart)
                          : reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(bstart) +
                                                         (((kBuckets + 1) * sizeof(PosT) + 15) & ~15));
    PosT *sorted = kSortedGlobal ? sorted_global + (long long)blockIdx.x * sorted_stride
                                 : reinterpret_cast<PosT *>(cnt + kCntWords) + (kSent ? 32 : 0);

    const int lane = threadIdx.x & 31;
    // (through REDUX: the compiler then knows the warp index -- and the segment bounds and
    // the parse position derived from it -- to be uniform and keeps them off the vector ALU)
    const int warp = LZ77_TOKLOOP == 2 ? (int)__reduce_max_sync(0xffffffffu, threadIdx.x >> 5)
                                       : (int)(threadIdx.x >> 5);
    const unsigned lt_mask = (1u << lane) - 1u;
    const int cnt_col = warp >> 1, cnt_sh = (warp & 1) * 16;
    static_assert(!kSortedGlobal && sizeof(PosT) == 2, "the token loop reads uint16 buckets from shared memory");
    const uint32_t sslot = smem_u32(cnt);
    uint32_t sdata = smem_u32(smem);             // shared-window addresses, computed once
    // (opaque: otherwise the compiler re-derives the shared window base from SR_CgaCtaId for
    // every token -- four instructions, one of them a special-register read, in front of the
    // token's first loads)
#ifndef LZ77_OPAQUE_SDATA
#define LZ77_OPAQUE_SDATA 1
#endif
    if (LZ77_TOKLOOP == 2 && LZ77_OPAQUE_SDATA) asm volatile("" : "+r"(sdata));
    const uint32_t sbstart = smem_u32(bstart);
    const uint32_t ssorted = smem_u32(sorted);
    const int sg = lane / kLanes, sl = lane % kLanes;  // group in the warp, lane in the group
    const int gshift = sg * kLanes;
    const unsigned gmask = kLanes == 32 ? 0xffffffffu : ((1u << (kLanes & 31)) - 1u) << gshift;

    if (threadIdx.x == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (kSent && threadIdx.x < 32) (sorted - 32)[threadIdx.x] = (PosT)0;  // the walk's first round may reach here
    __syncthreads();

    // ---- fused path: tiles parked in the ring, their place in the stream pending ----
    int seq = 0, done = 0;  // tiles parked / resolved by this CTA (uniform across the CTA)
    // Resolves parked tiles, oldest first, as far as their predecessors have reported their
    // token counts (decoupled look-back over the tiles: tiles are handed out by ticket, so
    // every predecessor is running or done); waits only when the ring is full or `drain`.
    auto resolve_parked = [&](bool drain) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        while (done < seq) {
            const bool must = drain || (seq - done) >= kRing;
            __syncthreads();  // orders the ring stores of this CTA before their reads
            if (warp == 0) {
                const long long gt = s_pend_tile[done % kRing];
                const int k_tile = s_pend_k[done % kRing];
                unsigned long long excl = 0;
                bool ok = true;
                if (gt > 0) {
                    long long idx = gt - 1 - lane;
                    while (true) {
                        unsigned long long sv = kLbInc;  // in front of the first tile: nothing
                        if (idx >= 0)
                            asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(sv) : "l"(F.status + idx) : "memory");
                        if (must) {
                            while (__any_sync(0xffffffffu, (sv >> 62) == 0)) {
                                if ((sv >> 62) == 0) {
                                    __nanosleep(64);
                                    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(sv) : "l"(F.status + idx) : "memory");
                                }
                            }
                        } else if (__any_sync(0xffffffffu, (sv >> 62) == 0)) {
                            ok = false;  // a predecessor is still parsing: try again after the next tile
                            break;
                        }
                        const unsigned inc_mask = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
                        unsigned long long v = sv & kLbVal;
                        if (inc_mask) {
                            const int first = __ffs(inc_mask) - 1;
                            v = lane <= first ? v : 0ull;
                        }
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                        excl += v;
                        if (inc_mask) break;
                        idx -= 32;
                    }
                }
                if (lane == 0) {
                    s_ok = ok ? 1 : 0;
                    if (ok) {
                        const unsigned long long incl = excl + (unsigned long long)k_tile;
                        if (gt == F.tile0 + n_tiles - 1) {  // the launch's last tile: running total
                            *F.total = incl;
                            if (F.host_total) {
                                *reinterpret_cast<volatile unsigned long long *>(F.host_total) = incl;
                                __threadfence_system();
                            }
                            __threadfence();
                        }
                        atomicExch(&F.status[gt], kLbInc | incl);
                        s_excl = excl;
                    }
                }
            }
            __syncthreads();
            if (!s_ok) return;
            // ring -> stream.  T is a multiple of 8, so the tile owns whole bytes of the
            // stream (nothing to zero or merge): single bytes up to the first aligned word,
            // 32-bit stores, single bytes.
            {
                const int k_tile = s_pend_k[done % kRing];
                const unsigned long long g0 = 4ull + 3ull * s_excl;  // byte offset in the stream
                const uint8_t *src =
                    F.ring + ((size_t)blockIdx.x * kRing + (size_t)(done % kRing)) * kRingSlot;
                uint8_t *dst = F.out + g0;
                const int nbytes = 3 * k_tile;
                int head = (int)((4ull - (g0 & 3ull)) & 3ull);
                if (head > nbytes) head = nbytes;
                for (int i = threadIdx.x; i < head; i += kWarps * 32) dst[i] = __ldcg(src + i);
                const int nwords = (nbytes - head) >> 2;
                const uint32_t *sw = reinterpret_cast<const uint32_t *>(src);
                uint32_t *dw = reinterpret_cast<uint32_t *>(dst + head);
                const int bsh = head * 8;
                for (int w = threadIdx.x; w < nwords; w += kWarps * 32)
                    dw[w] = __funnelshift_r(__ldcg(sw + w), __ldcg(sw + w + 1), bsh);
                for (int i = head + 4 * nwords + (int)threadIdx.x; i < nbytes; i += kWarps * 32)
                    dst[i] = __ldcg(src + i);
                if (s_pend_tile[done % kRing] == 0 && F.write_header && threadIdx.x == 0)
                    *reinterpret_cast<uint32_t *>(F.out) = (uint32_t)P.sb | ((uint32_t)P.la << 16);
            }
            done++;
        }
    };

    uint32_t phase = 0;
    for (long long tile_i = blockIdx.x;; tile_i += gridDim.x, phase ^= 1u) {
        if (kFused) {
            // tiles are handed out in order: whoever looks back at a tile finds it running
            if (threadIdx.x == 0) {
                const unsigned int t = atomicAdd(F.ticket, 1u);
                // every CTA draws until it is turned away: the last draw of the launch
                // leaves the ticket at zero for the next launch that uses this slot
                if ((long long)t == n_tiles + (long long)gridDim.x - 1) atomicExch(F.ticket, 0u);
                s_tile = (long long)t;
            }
            __syncthreads();
            tile_i = s_tile;
        }
        if (tile_i >= n_tiles) break;
        const long long tile_lo = tile_i * tile_bytes;
        // the window starts at the block (independent blocks) or, in history mode, reaches
        // back across block seams into the `pre` valid bytes in front of `in`
        const long long blk_lo = P.history ? -pre : (tile_lo >> P.block_shift) << P.block_shift;

        // ---- stage history + tile with one TMA bulk copy -------------------
        long long hist = tile_lo - blk_lo;
        if (hist > P.window) hist = P.window;
        const int hist_al = (int)((hist + 15) & ~15LL);
        const long long src_lo = tile_lo - hist_al;
        long long src_hi = tile_lo + tile_bytes;
        if (src_hi > n) src_hi = n;
        const int bytes = (int)(src_hi - src_lo);
        const int bulk = bytes & ~15;
        const int dst0 = hist_cap - hist_al;  // smem index of global byte src_lo

        if (threadIdx.x == 0 && bulk > 0) {
            fence_proxy_async();
            mbar_expect_tx(&mbar, (uint32_t)bulk);
            tma_load_1d(smem + dst0, in + src_lo, (uint32_t)bulk, &mbar);
        }
        for (int i = bulk + threadIdx.x; i < bytes + 64; i += kThreads)
            smem[dst0 + i] = (i < bytes) ? in[src_lo + i] : (uint8_t)0;
        for (int i = threadIdx.x; i < kCntWords; i += kThreads) cnt[i] = 0u;
        if (bulk > 0) mbar_wait(&mbar, phase);
        __syncthreads();

        // ---- build: stable counting sort of the positions by key -----------
        const int rows = (bytes + kThreads - 1) / kThreads;  // rows of 32 positions per warp
        const int cbase = dst0 + warp * rows * 32;
        const int cend = min(dst0 + bytes, cbase + rows * 32);
        // Counters: one 16-bit count per (bucket, warp); the word of (column, key) is
        // cnt[col * kBuckets + key] with the even warp of the column in its low half -- the keys
        // of a row spread over all banks (6.19 -> 6.00 ms against key-major words), and the scan
        // below reads four keys per 128-bit load.  Measured and not kept: no atomics at all -- a
        // warp owns its counters, so the first lane of every group of equal keys (a second
        // MATCH.ANY per row) can add with a plain load and store -- 7.36 ms: MATCH.ANY, about 55
        // cycles of its unit per row, is the build's bottleneck, not ATOMS; ranks from eleven
        // ballots instead of MATCH.ANY build faster alone (2.2 vs 2.9 ms) but take issue slots
        // from the tiles that are parsing: 6.42 ms.
        for (int i = cbase + lane; i < cend; i += 32)
            atomicAdd(&cnt[cnt_col * kBuckets + bucket_key(smem, i)], 1u << cnt_sh);
        __syncthreads();
        {
            // thread t owns buckets [t*per, (t+1)*per): exclusive scan across the
            // warps' counters, then across buckets
            constexpr int per = kBuckets / kThreads;
            static_assert(per == 4, "four buckets per thread: one 128-bit load per counter column");
            uint32_t tot[per] = {0u, 0u, 0u, 0u};
            uint32_t x[kWarps / 2][per];  // exclusive offsets of the column's two warps, packed
#pragma unroll
            for (int w = 0; w < kWarps / 2; w++) {
                const uint4 v = *reinterpret_cast<const uint4 *>(cnt + w * kBuckets + threadIdx.x * per);
                const uint32_t vv[per] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int b = 0; b < per; b++) {
                    const uint32_t lo = vv[b] & 0xffffu, hi = vv[b] >> 16;
                    x[w][b] = tot[b] | ((tot[b] + lo) << 16);
                    tot[b] += lo + hi;
                }
            }
            const uint32_t sum = tot[0] + tot[1] + tot[2] + tot[3];
            uint32_t base = block_exclusive_scan_u32<kThreads>(sum, s_warp, &s_total);
            uint32_t first[per];
#pragma unroll
            for (int b = 0; b < per; b++) {
                if (kSent) {
                    // bucket k: sentinel at base + k, entries behind it; the start is folded
                    // into the counters of all warps, so the scatter reads nothing else
                    first[b] = base + (uint32_t)(threadIdx.x * per + b) + 1u;
                    sorted[first[b] - 1u] = (PosT)0;
                } else {
                    first[b] = 0u;
                    bstart[threadIdx.x * per + b] = (PosT)base;
                }
                base += tot[b];
            }
            if (!kSent && threadIdx.x == kThreads - 1) bstart[kBuckets] = (PosT)base;
#pragma unroll
            for (int w = 0; w < kWarps / 2; w++) {
                uint4 v;
                v.x = x[w][0] + first[0] * 0x10001u, v.y = x[w][1] + first[1] * 0x10001u;
                v.z = x[w][2] + first[2] * 0x10001u, v.w = x[w][3] + first[3] * 0x10001u;
                *reinterpret_cast<uint4 *>(cnt + w * kBuckets + threadIdx.x * per) = v;
            }
        }
        __syncthreads();
        // Scatter, row by row.  Every lane takes part -- a lane past the end with a key of its
        // own -- because a shuffle under a partial mask costs a second MATCH.ANY.  (kBatch rows
        // can be issued together; one at a time is fastest.)
        constexpr int kBatch = LZ77_SCATTER_BATCH;
        for (int r0 = 0; r0 < rows; r0 += kBatch) {
            int key[kBatch];
            unsigned peers[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; b++) {
                const int i = cbase + (r0 + b) * 32 + lane;
                key[b] = (r0 + b < rows && i < cend) ? bucket_key(smem, i) : kBuckets + lane;
            }
#pragma unroll
            for (int b = 0; b < kBatch; b++) {
#if LZ77_BALLOT_RANK
                unsigned m = 0xffffffffu;
#pragma unroll
                for (int bit = 0; bit < 2 * kKeyBits + 1; bit++) {
                    const bool one = (key[b] >> bit) & 1;
                    const unsigned bal = __ballot_sync(0xffffffffu, one);
                    m &= one ? bal : ~bal;
                }
                // (lanes past the end carry bit 2*kKeyBits and differ among themselves)
                peers[b] = key[b] < kBuckets ? m : 1u << lane;
#else
                peers[b] = __match_any_sync(0xffffffffu, key[b]);
#endif
            }
#pragma unroll
            for (int b = 0; b < kBatch; b++) {
                const int i = cbase + (r0 + b) * 32 + lane;
                const bool valid = key[b] < kBuckets;
                const int leader = __ffs(peers[b]) - 1;
                uint32_t old = 0;
                if (valid && lane == leader)
                    old = (atomicAdd(&cnt[cnt_col * kBuckets + key[b]], (uint32_t)__popc(peers[b]) << cnt_sh) >>
                           cnt_sh) & 0xffffu;
                old = __shfl_sync(0xffffffffu, old, leader);
                if (valid) {
                    const int slot = (kSent ? 0 : (int)bstart[key[b]]) + (int)old + __popc(peers[b] & lt_mask);
                    sorted[slot] = (PosT)i;
                }
            }
        }
        if (kSortedGlobal) __threadfence_block();
        __syncthreads();

        const int tile_idx = dst0 + hist_al;  // shared-memory index of the tile's first byte
        if (kBackWalk) {
            // own index of every tile position in its bucket list; the counters are dead
            // now and their area holds the table (tile_bytes * 2 bytes)
            uint16_t *slot_of = reinterpret_cast<uint16_t *>(cnt);
            for (int e = threadIdx.x; e < bytes + (kSent ? kBuckets : 0); e += kThreads) {
                const int q = (int)sorted[e];  // (a sentinel reads as position 0: in front of the tile)
                if (q >= tile_idx) slot_of[q - tile_idx] = (uint16_t)(kSent ? 2 * e : e);  // (kSent: byte offset)
            }
            __syncthreads();
        }

        // ---- parse: one group of kLanes lanes per segment -------------------
        const long long seg_lo =
            tile_lo + (long long)(warp * kSegsPerWarp + sg) * kSeg;
        const long long sgm = seg_lo / kSeg;  // global segment index
        if (seg_lo < n && !LZ77_BUILD_ONLY) {
            long long seg_hi = seg_lo + kSeg;
            if (seg_hi > n) seg_hi = n;
            const int seg_end = (int)(seg_hi - src_lo) + dst0;
            int p0 = (int)(seg_lo - src_lo) + dst0;
            const int first_idx = dst0 + hist_al - (int)hist;  // oldest byte a match may start at
            uint32_t *tok_row = kFused ? nullptr : tok_tmp + sgm * kSeg;
            const int len_shift = P.ob, lit_shift = P.ob + P.lb;
            const int la = P.la, window = P.window;
            uint32_t *tok_at = tok_row;  // a running pointer: no address arithmetic per token
            // fused: the segment's tokens stay in shared memory (the counters of the build)
            uint32_t tok_sa = smem_u32(cnt) + (uint32_t)warp * (kTokBuf * 4u);
            const uint32_t tok_sa_end = tok_sa + kTokBuf * 4u;
            uint32_t *spill_at = kFused ? F.spill + ((size_t)blockIdx.x * kWarps + warp) * kTokSpill
                                        : nullptr;
            int ntok_f = 0;


            if constexpr (kBackWalk && LZ77_TOKLOOP == 2) {
                // ---- token loop written against the ALU pipe (ncu: the vector ALU is the
                // busiest unit of this kernel, 76 % of its cycles, the FMA pipe idles) ----
                // The running best of a lane is ONE packed key, len * 65536 - start: a single
                // multiply-add (FMA pipe) and a max per round replace two compares and two
                // selects, equal lengths prefer the older start for free, and the REDUX result
                // unpacks on the uniform datapath together with the loop control.  The last
                // byte of a segment (no match possible, tree.c:136) leaves the loop, so no
                // test of the lookahead length per token; a position without reach finds no
                // candidate inside its window and falls through as a literal.
                constexpr int kNone = -65535;  // length 0, no start
                constexpr int kFar = 1 << 30;  // a start whose key loses against kNone
                const int la1 = la - 1;
                const int last = seg_end - 1;
                uint32_t len_mul = 1u << len_shift, lit_mul = 1u << lit_shift;
                asm("" : "+r"(len_mul), "+r"(lit_mul));  // (opaque: keeps the multiplies)
#if LZ77_DEFER_STORE
                uint32_t pend_lit = 0u, pend_rest = 0u;
#endif
                const uint32_t slot_base = sslot - 2u * (uint32_t)tile_idx;   // slot table by staged index
                const uint32_t lane_base = ssorted - 2u - 2u * (uint32_t)lane;
                int ntok = 0;
                while (p0 < last) {
                    const int max_len = min(la1, last - p0);          // lz77.c:87,134 + tree.c:136
                    const int lo_idx = max(p0 - window, first_idx);   // lz77.c:101-105
                    const uint32_t w = sdata + (uint32_t)(p0 & ~3);
                    const int sh = p0 * 8;  // (the funnel shift wraps: only bits 3..4 count)
                    const uint32_t a0 = lds32(w), a1 = lds32(w + 4), a2 = lds32(w + 8);
                    const uint32_t tgt0 = __funnelshift_r(a0, a1, sh);
                    const uint32_t tgt1 = __funnelshift_r(a1, a2, sh);
                    // own place in the bucket list, as a byte offset into the sorted array
                    const uint32_t c_hi2 = lds16(slot_base + 2u * (uint32_t)p0);
#if LZ77_DEFER_STORE
                    // the previous token leaves here, behind this token's first loads: its
                    // literal (a shared-memory load at the end of the last pass) has arrived
                    if (ntok > 0 && lane == 0) tok_row[ntok - 1] = pend_lit * lit_mul + pend_rest;
#endif
                    int best = kNone;
                    uint32_t ca = lane_base + c_hi2;  // this lane's entry: own place - 1 - lane
                    bool fwd = false;
                    while (true) {
                        // the zero entry in front of the bucket ends the walk like an entry that
                        // has left the window; lanes beyond it read entries of other buckets,
                        // which differ within the first two bytes -- a length below 2 never
                        // reaches the token (the scan below decides those)
                        const int q = (int)lds16(ca);
                        const bool in = q >= lo_idx;  // (q = 0 lies in front of every window)
                        // (decided in front of the compare, while the predicate is at hand: an
                        // entry outside the window gets a start that can never win)
                        const bool left = __any_sync(0xffffffffu, !in);
                        const int qk = in ? q : kFar;
                        const int l = min(cand_match_len(sdata, q, in, p0, tgt0, tgt1), max_len);
                        best = max(best, l * 65536 - qk);
                        ca -= 64u;
                        if (left) break;  // left the window (or the bucket)
                        if (__all_sync(0xffffffffu, l >= max_len)) {
                            fwd = true;
                            break;
                        }
                    }
                    if (fwd) {
                        // every candidate of a full round matched to the maximum (runs, short
                        // periods): the oldest such entry of the window ends the search
                        // first entry inside the window.  Sentinel layout: entries are distinct
                        // ascending positions, so it is not more than `reach` places in front of
                        // the own one; a probe that lands in another bucket (or on the sentinel)
                        // fails the key test and counts as "in front of the window"
                        const int c_hi = (int)(c_hi2 >> 1);
                        int lo = max(c_hi - (p0 - lo_idx), 0), hi = c_hi;
                        const int key = pair_key(tgt0, tgt0 >> 8);
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            const int qm = (int)lds16(ssorted + 2u * mid);
                            if (qm < lo_idx || bucket_key(smem, qm) != key)
                                lo = mid + 1;
                            else
                                hi = mid;
                        }
                        best = kNone;
                        for (int i = lo; i < c_hi; i += 32) {
                            const int idx = i + lane;
                            const bool in = idx < c_hi;
                            const int q = in ? (int)lds16(ssorted + 2u * idx) : 0;
                            const int l = min(cand_match_len(sdata, q, in, p0, tgt0, tgt1), max_len);
                            best = max(best, in ? l * 65536 - q : kNone);
                            if (__any_sync(0xffffffffu, best >= max_len * 65536 - 65535)) break;
                        }
                    }
                    const int kbest = __reduce_max_sync(0xffffffffu, best);
                    int len = (kbest + 65535) >> 16;
                    int off = p0 + kbest - (len << 16);  // p0 - start
                    if (len < 2) {
                        // length 1 (such bytes sit in many buckets): scan of the staged window
                        const int q1 = oldest_byte_match(sdata, lo_idx, p0, tgt0 & 0xffu, lane);
                        len = q1 >= 0 ? 1 : 0;
                        off = q1 >= 0 ? p0 - q1 : 0;
                    }
                    // the fields do not overlap: two multiply-adds (FMA pipe) instead of
                    // two shifts and an OR
#if LZ77_DEFER_STORE
                    pend_lit = lds8(sdata + (uint32_t)(p0 + len));
                    pend_rest = (uint32_t)len * len_mul + (uint32_t)off;
#else
                    const uint32_t lit = lds8(sdata + (uint32_t)(p0 + len));
                    const uint32_t tok = lit * lit_mul + ((uint32_t)len * len_mul + (uint32_t)off);
                    if (lane == 0) tok_row[ntok] = tok;
#endif
                    ntok++;
                    p0 += len + 1;
                }
#if LZ77_DEFER_STORE
                if (ntok > 0 && lane == 0) tok_row[ntok - 1] = pend_lit * lit_mul + pend_rest;
#endif
                if (p0 == last) {  // the segment's last byte: a literal
                    if (lane == 0) tok_row[ntok] = lds8(sdata + (uint32_t)p0) << lit_shift;
                    ntok++;
                    p0++;
                }
                tok_at = tok_row + ntok;
            }
            while (!(kBackWalk && LZ77_TOKLOOP == 2) && p0 < seg_end) {
                const int max_len = min(la, seg_end - p0) - 1;  // lz77.c:87,134 + tree.c:136
                const int reach = min(p0 - first_idx, window);  // lz77.c:101-105
                int len = 0, off = 0;

                if (max_len > 0 && reach > 0) {
                    const int lo_idx = p0 - reach;
                    // lookahead: 16 bytes at p0 from five aligned words
                    uint32_t tgt[4];
                    {
                        const uint32_t w = sdata + (uint32_t)(p0 & ~3);
                        const int sh = (p0 & 3) * 8;
                        const uint32_t a0 = lds32(w), a1 = lds32(w + 4), a2 = lds32(w + 8);
                        tgt[0] = __funnelshift_r(a0, a1, sh);
                        tgt[1] = __funnelshift_r(a1, a2, sh);
                        if (!kSmallLA || !LZ77_LAZY_TGT) {  // (LA <= 16 reads bytes 8..15 on demand)
                            const uint32_t a3 = lds32(w + 12), a4 = lds32(w + 16);
                            tgt[2] = __funnelshift_r(a2, a3, sh);
                            tgt[3] = __funnelshift_r(a3, a4, sh);
                        } else {
                            tgt[2] = tgt[3] = 0u;
                        }
                    }
                    const int key = pair_key(tgt[0], tgt[0] >> 8);
                    int best_len = 0, best_q = 0;
                    if (kBackWalk) {
                        // Candidates are walked BACKWARDS from the position's own place in its
                        // bucket list (slot table built after the sort): no bucket size, no lower
                        // bound of the window start, no round spent on entries that have left the
                        // window.  The oldest of the longest still wins: ties go to the candidate
                        // seen later, which is the older one.  A round that matches to the maximum
                        // length in every lane (runs, short periods: the whole window would
                        // follow) switches to the forward walk from the window's oldest entry,
                        // which ends at the first maximum-length match.
                        const int c_lo = (int)lds16(sbstart + 2u * key);
                        const int c_hi = (int)lds16(sslot + 2u * (uint32_t)(p0 - tile_idx));
                        int ci = c_hi - 1;
                        bool fwd = false;
                        while (true) {
                            const int idx = ci - sl;
                            const int q = idx >= c_lo ? (int)lds16(ssorted + 2u * idx) : -1;
                            const bool in = q >= lo_idx;  // (entries in front of the own one are older)
                            const int l = round_match_len(sdata, q, in, p0, tgt[0], tgt[1], tgt[2], tgt[3],
                                                          max_len);
                            if (l > 0 && l >= best_len) {
                                best_len = l;
                                best_q = q;
                            }
                            ci -= kLanes;
                            if (__any_sync(gmask, !in)) break;  // left the window (or the bucket)
                            if (__all_sync(gmask, l >= max_len)) {
                                fwd = true;
                                break;
                            }
                        }
                        if (fwd) {
                            int lo = c_lo, hi = c_hi;  // first entry inside the window
                            while (lo < hi) {
                                const int mid = (lo + hi) >> 1;
                                if ((int)lds16(ssorted + 2u * mid) < lo_idx)
                                    lo = mid + 1;
                                else
                                    hi = mid;
                            }
                            best_len = 0;
                            for (int i = lo; i < c_hi; i += kLanes) {
                                const int idx = i + sl;
                                const bool in = idx < c_hi;
                                const int q = in ? (int)lds16(ssorted + 2u * idx) : p0;
                                const int l = round_match_len(sdata, q, in, p0, tgt[0], tgt[1], tgt[2],
                                                              tgt[3], max_len);
                                if (l > best_len) {
                                    best_len = l;
                                    best_q = q;
                                }
                                if (__any_sync(gmask, best_len >= max_len)) break;
                            }
                        }
                    } else {

                        const int bs = (int)lds16(sbstart + 2u * key);
                        const int bn = (int)lds16(sbstart + 2u * key + 2u) - bs;
                        const uint32_t se = ssorted + 2u * bs;  // shared address of the bucket
                        // candidates: bucket entries in [lo_idx, p0), oldest first; short
                        // buckets are walked from their start, long ones from the window's
                        // lower bound
                        int i = bn <= kLinearScan
                                    ? 0
                                    : group_lower_bound_s<kLanes>(se, bn, lo_idx, sl, gmask, gshift);
                        for (; i < bn; i += kLanes) {
                            const int idx = i + sl;
                            // LA <= 16: the last round of a bucket reads on into the next one (or the
                            // padding behind the list).  Whatever turns up there is only accepted as
                            // an in-window position whose bytes really match, and such a position with
                            // two or more matching bytes is in this bucket anyway; a one-byte match is
                            // rescanned below.  Saves the bounds predicate in every round.
                            const int q = (kSmallLA || idx < bn) ? (int)lds16(se + 2u * idx) : 0x7fffffff;
                            const bool in = q >= lo_idx && q < p0;
                            if (kSmallLA) {
                                const int l = round_match_len(sdata, q, in, p0, tgt[0], tgt[1], tgt[2], tgt[3],
                                                              max_len);
                                // nearer than anything this lane has seen: must be longer
                                if (l > best_len) {
                                    best_len = l;
                                    best_q = q;
                                }
                            } else if (in) {
                                const int l = match_len_long(sdata, q, p0, tgt, max_len);
                                if (l > best_len) {
                                    best_len = l;
                                    best_q = q;
                                }
                            }
                            if (__any_sync(gmask, best_len >= max_len || q >= p0)) break;
                        }
                    }
                    // (no candidate: length 0 in the top bits, the start is not used)
                    const uint32_t k = __reduce_max_sync(
                        gmask, ((uint32_t)best_len << 20) | (0xfffffu - (uint32_t)best_q));
                    len = (int)(k >> 20);
                    int q_best = (int)(0xfffffu - (k & 0xfffffu));
                    if (len < 2) {
                        // length 1: the oldest byte of the window equal to the first
                        // lookahead byte -- forward SWAR scan of the staged window, 512
                        // bytes per step (such bytes sit in many buckets)
                        const uint32_t b0x4 = (tgt[0] & 0xffu) * 0x01010101u;
                        int q1 = 0x7fffffff;
                        for (int base = lo_idx & ~15; base < p0; base += kLanes * 16) {
                            const int g = base + sl * 16;
                            if (g < p0) {
                                const uint32_t ga = sdata + (uint32_t)g;
                                uint32_t wv[4];
                                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(wv[0]), "=r"(wv[1]), "=r"(wv[2]), "=r"(wv[3])
                                             : "r"(ga));
#pragma unroll
                                for (int wi = 3; wi >= 0; wi--) {
                                    uint32_t m = zero_bytes(wv[wi] ^ b0x4);
                                    while (m) {
                                        const int bit = __ffs(m) - 1;
                                        m ^= 1u << bit;
                                        const int q = g + 4 * wi + (bit >> 3);
                                        if (q >= lo_idx && q < p0 && q < q1) q1 = q;
                                    }
                                }
                            }
                            if (__any_sync(gmask, q1 != 0x7fffffff)) break;
                        }
                        q1 = (int)__reduce_min_sync(gmask, (unsigned)q1);
                        len = q1 != 0x7fffffff ? 1 : 0;
                        q_best = q1;
                    }
                    off = len ? p0 - q_best : 0;
                }

                const uint32_t lit = lds8(sdata + (uint32_t)(p0 + len));
                const uint32_t tok =
                    (uint32_t)off | ((uint32_t)len << len_shift) | (lit << lit_shift);
                // one 4-byte store per token by one lane: fewer instructions than
                // collecting rows of 32 tokens in registers (measured 9.6 -> 9.4 ms), and
                // L2 merges the sectors
                if (kFused) {
                    if (tok_sa < tok_sa_end) {
                        if (sl == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(tok_sa), "r"(tok));
                        tok_sa += 4u;
                    } else {
                        if (sl == 0) *spill_at = tok;
                        spill_at++;
                    }
                    ntok_f++;
                } else {
                    if (sl == 0) *tok_at = tok;
                    tok_at++;
                }
                p0 += len + 1;
            }
            if (kFused) {
                if (lane == 0) s_cnt[warp] = ntok_f;
            } else {
                const int ntok = (int)(tok_at - tok_row);
                if (sl == 0) seg_ntok[sgm] = (uint32_t)ntok;
            }
        } else if (kFused) {
            if (lane == 0) s_cnt[warp] = 0;
        }
        __syncthreads();  // the next tile overwrites the staged data and the buckets

        if (kFused) {
            // ---- pack: 3 bytes per token into shared memory (the bucket lists are dead) ----
            int pre_w = 0, k_tile = 0;
#pragma unroll
            for (int w = 0; w < kWarps; w++) {
                const int c = s_cnt[w];
                if (w < warp) pre_w += c;
                k_tile += c;
            }
            uint8_t *pk = reinterpret_cast<uint8_t *>(sorted);
            {
                const uint32_t *buf = cnt + warp * kTokBuf;
                const uint32_t *sp = F.spill + ((size_t)blockIdx.x * kWarps + warp) * kTokSpill;
                const int mine = s_cnt[warp];
                for (int t = lane; t < mine; t += 32) {
                    const uint32_t tok = t < kTokBuf ? buf[t] : sp[t - kTokBuf];
                    uint8_t *b = pk + 3 * (pre_w + t);
                    b[0] = (uint8_t)tok;
                    b[1] = (uint8_t)(tok >> 8);
                    b[2] = (uint8_t)(tok >> 16);
                }
            }
            __syncthreads();
            // ---- park the packed tile in this CTA's ring (global memory, L2-resident) and
            //      publish its token count; its place in the stream is resolved later, so a
            //      slow tile ahead does not stall this CTA (no convoy behind the slowest tile) ----
            {
                uint4 *slot = reinterpret_cast<uint4 *>(
                    F.ring + ((size_t)blockIdx.x * kRing + (size_t)(seq % kRing)) * kRingSlot);
                const int nchunks = (3 * k_tile + 15) >> 4;
                for (int c = threadIdx.x; c < nchunks; c += kThreads)
                    slot[c] = reinterpret_cast<const uint4 *>(pk)[c];
                const long long gt = F.tile0 + tile_i;  // index in the whole call
                if (threadIdx.x == 0) {
                    s_pend_tile[seq % kRing] = gt;
                    s_pend_k[seq % kRing] = k_tile;
                    if (gt > 0) atomicExch(&F.status[gt], kLbAgg | (unsigned long long)k_tile);
                }
                seq++;
            }
            // (the barrier at the top of resolve orders the ring stores before their reads)
        }
        if (kFused) resolve_parked(false);
    }
    if (kFused) resolve_parked(true);  // no tiles left to draw: wait for the parked ones
}

// ---------------------------------------------------------------------------


bool parse_bucket_fused(const Params &P)
{
    return P.fused_pack && P.tbits == 24 && P.window <= 8191;
}

// scratch of the fused path for a call over n_total input bytes: look-back words of all
// its tiles, the tickets (one per launch that may be in flight), the spill area of one
// resident grid
constexpr int kMaxTickets = 64;
constexpr int kFusedGridMax = 160 * 4;
static size_t fused_status_bytes(long long n_tiles)  // (the areas behind it stay 256-byte aligned)
{
    return ((size_t)(n_tiles + 64) * 8 + 255) & ~(size_t)255;
}

size_t parse_bucket_fused_scratch(long long n_total)
{
    const long long tile_bytes = (long long)LZ77_PARSE_WARPS * kSegBytes;
    const long long n_tiles = (n_total + tile_bytes - 1) / tile_bytes;
    // (two sets of spill rows and ring slots: the launches of a chunked call alternate
    // between two streams and may run side by side)
    const size_t ctas = (size_t)(n_tiles < kFusedGridMax ? (n_tiles > 0 ? n_tiles : 1) : kFusedGridMax);
    return fused_status_bytes(n_tiles) + (size_t)kMaxTickets * 4 + 256 +
           2 * (ctas * LZ77_PARSE_WARPS * kTokSpill * 4 + ctas * kRing * kRingSlot) + 1024;
}

// Bytes after which the greedy parse restarts: part of the stream's specification
// (lz77_gpu_segment_size()).  One value today; it is a function of the parameters because a
// parser that runs several shorter segments per warp was measured (see the note at kLanes).
int parse_segment_bytes(int window, int la, bool fused_pack)
{
    (void)window, (void)la, (void)fused_pack;
    return kSegBytes;
}

cudaError_t launch_parse_bucket(const uint8_t *d_in, long long n_in, long long pre,
                                const Params &P, uint32_t *tok_tmp, uint32_t *seg_ntok,
                                cudaStream_t st)
{
    constexpr int kW = LZ77_PARSE_WARPS, kL = LZ77_PARSE_LANES;
    const bool small_la = P.la <= 16;
    // (strictly more than the window: staged position 0 -- the token loop's "no candidate"
    // value and the bucket lists' sentinel -- then lies in front of every window, also when
    // the window is a multiple of 16 bytes and for SB = 1, whose usable window is empty)
    const int hist_cap = (P.window & ~15) + 16;
    const long long tile_bytes = (long long)kW * (32 / kL) * kSegBytes;
    const long long n_tiles = (n_in + tile_bytes - 1) / tile_bytes;
    if (n_tiles == 0) return cudaSuccess;
    const size_t data_cap = (size_t)hist_cap + (size_t)tile_bytes + 64;
    size_t smem = (data_cap + 15) & ~(size_t)15;               // staged bytes
    const bool sentinel = small_la && kL == 32 && LZ77_BACKWALK != 0 && LZ77_TOKLOOP == 2;
    if (!sentinel) smem += ((kBuckets + 1) * sizeof(uint16_t) + 15) & ~(size_t)15;  // bucket starts
    smem += (size_t)kBuckets * (kW / 2) * 4;                   // per-warp counters
    smem += data_cap * sizeof(uint16_t) + 80;                  // sorted positions + one round of padding
    if (sentinel) smem += (32 + kBuckets) * sizeof(uint16_t);  // front padding + one zero entry per bucket
    auto kern = small_la ? lz77_parse_bucket_kernel<true, kW, kL, uint16_t, false, false, kSegBytes>
                         : lz77_parse_bucket_kernel<false, kW, kL, uint16_t, false, false, kSegBytes>;
    cudaError_t rc =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    kern<<<(unsigned)n_tiles, kW * 32, smem, st>>>(d_in, n_in, pre, P, hist_cap, n_tiles, tok_tmp,
                                                   seg_ntok, (uint16_t *)nullptr, 0, FusedEmit{});
    return cudaGetLastError();
}

// Zeroes the look-back words and the tickets of a call over n_total bytes.  Launches of one
// call may run on several streams: the caller orders this in front of all of them.
cudaError_t launch_parse_bucket_fused_reset(long long n_total, void *scratch, cudaStream_t st)
{
    const long long tile_bytes = (long long)LZ77_PARSE_WARPS * kSegBytes;
    const long long n_tiles_total = (n_total + tile_bytes - 1) / tile_bytes;
    return cudaMemsetAsync(scratch, 0, fused_status_bytes(n_tiles_total) + (size_t)kMaxTickets * 4, st);
}

// The fused path: search + parse + pack in one persistent kernel.  `scratch` as sized by
// parse_bucket_fused_scratch() for the whole call; `first`: the launch that holds the start
// of the input (writes the header); `reset`: zero the look-back state first (single-launch
// calls; chunked calls reset once, in front of all their streams); slot = the ticket of
// this launch.
cudaError_t launch_parse_bucket_fused(const uint8_t *d_in, long long lo, long long n_in,
                                      long long n_total, long long pre, bool first, bool reset,
                                      int slot, const Params &P, void *scratch, uint8_t *d_out,
                                      unsigned long long *total, unsigned long long *host_total,
                                      cudaStream_t st)
{
    constexpr int kW = LZ77_PARSE_WARPS;
    const bool small_la = P.la <= 16;
    const int hist_cap = (P.window + 15) & ~15;
    const long long tile_bytes = (long long)kW * kSegBytes;
    const long long n_tiles_total = (n_total + tile_bytes - 1) / tile_bytes;
    const long long n_tiles = (n_in + tile_bytes - 1) / tile_bytes;
    char *p = (char *)scratch;
    unsigned long long *status = (unsigned long long *)p;
    p += fused_status_bytes(n_tiles_total);
    unsigned int *tickets = (unsigned int *)p;
    p += (size_t)kMaxTickets * 4 + 256;
    const size_t ctas = (size_t)(n_tiles_total < kFusedGridMax ? (n_tiles_total > 0 ? n_tiles_total : 1)
                                                               : kFusedGridMax);
    const size_t spill_bytes = ctas * kW * kTokSpill * 4;
    const size_t ring_bytes = ctas * kRing * kRingSlot;
    p += (size_t)(slot & 1) * (spill_bytes + ring_bytes);  // the set of this launch's stream
    uint32_t *spill = (uint32_t *)p;
    p += spill_bytes;
    uint8_t *ring = (uint8_t *)p;
    cudaError_t rc;
    if (reset) {
        rc = launch_parse_bucket_fused_reset(n_total, scratch, st);
        if (rc != cudaSuccess) return rc;
    }
    if (slot < 0 || slot >= kMaxTickets) return cudaErrorInvalidValue;
    if (n_tiles == 0) {  // empty input: the header alone
        if (first) {
            const uint32_t hdr = (uint32_t)P.sb | ((uint32_t)P.la << 16);
            rc = cudaMemcpyAsync(d_out, &hdr, 4, cudaMemcpyHostToDevice, st);
            if (rc != cudaSuccess) return rc;
            rc = cudaMemsetAsync(total, 0, 8, st);
            if (rc != cudaSuccess) return rc;
        }
        return cudaSuccess;
    }
    const size_t data_cap = (size_t)hist_cap + (size_t)tile_bytes + 64;
    size_t smem = (data_cap + 15) & ~(size_t)15;
    smem += ((kBuckets + 1) * sizeof(uint16_t) + 15) & ~(size_t)15;
    smem += (size_t)kBuckets * (kW / 2) * 4;   // counters, then the tokens of the 8 segments
    static_assert((size_t)kBuckets * (LZ77_PARSE_WARPS / 2) * 4 >= (size_t)LZ77_PARSE_WARPS * kTokBuf * 4,
                  "the token buffers live in the counter area");
    // sorted positions, then the tile's packed tokens (worst case one token per byte)
    size_t tail = data_cap * sizeof(uint16_t) + 80;
    const size_t pk_need = (size_t)tile_bytes * 3 + 32;
    if (tail < pk_need) tail = pk_need;
    smem += tail;
    auto kern = small_la ? lz77_parse_bucket_kernel<true, kW, 32, uint16_t, false, true, kSegBytes>
                         : lz77_parse_bucket_kernel<false, kW, 32, uint16_t, false, true, kSegBytes>;
    rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    rc = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kW * 32, smem);
    if (rc != cudaSuccess) return rc;
    long long grid = (long long)sms * (per_sm > 0 ? per_sm : 1);
    if (grid > kFusedGridMax) grid = kFusedGridMax;
    if (grid > n_tiles) grid = n_tiles;
    FusedEmit F;
    F.out = d_out;
    F.status = status;
    F.ticket = tickets + slot;
    F.total = total;
    F.host_total = host_total;
    F.spill = spill;
    F.ring = ring;
    F.tile0 = lo / tile_bytes;
    F.write_header = first ? 1 : 0;
    kern<<<(unsigned)grid, kW * 32, smem, st>>>(d_in + lo, n_in, pre, P, hist_cap, n_tiles, nullptr,
                                                nullptr, (uint16_t *)nullptr, 0, F);
    return cudaGetLastError();
}

}  // namespace lz77
