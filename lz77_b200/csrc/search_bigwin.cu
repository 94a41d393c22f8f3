// search_bigwin.cu -- bucketed longest-match search for large windows
// (SB > 8191, e.g. -s 65535 -l 255), sm_100a.
//
// The per-tile bucket build of search_bucket.cu does not fit shared memory when
// the window is 64 KiB, so the buckets are built once per independent block
// (512 KiB) into HBM / L2 and shared by all the tiles of the block:
//
//   lz77_block_sort_kernel   one CTA per block (512 KiB): its positions are sorted by
//       key = (x[q] & 127) << 6 | x[q+1] & 63 (8192 buckets) with a stable two-pass
//       LSD radix sort (digit x[q+1]&63, then digit x[q]&127); per-warp digit counters
//       + MATCH.ANY ranks keep every bucket in ascending position order.
//       Output: sorted positions (uint32) and 8193 bucket starts per block.
//   lz77_parse_bigwin_kernel one CTA (16 warps) per 16 KiB tile: TMA-stages up to
//       64 KiB of history + the tile, then warp w parses segment w exactly like
//       search_bucket.cu, reading its candidates from the block's bucket lists:
//       warp-ary lower bound of the window start, 32 candidates per round
//       verified against shared memory, REDUX of (length, oldest start).
//       A length-1 match (any bucket of the first byte) is found by a forward
//       SWAR scan of the staged window from its oldest byte.
//
// Same result as the exhaustive scan in encode.cu (tests compare both against
// the oracle byte for byte).
#include "kernels.cuh"
#include "match.cuh"

namespace lz77 {

// key = low 7 bits of x[q], low 6 bits of x[q+1] (8192 buckets per block): a perfect hash of
// the byte pair for ASCII text, an even spread for binary data.  Measured with the byte pair
// itself as the key (8 + 8 bits, 65536 buckets: one candidate round per token on random data
// instead of two or three and a lower-bound search): 35.8 vs 36.0 ms on 256 MiB -- the parse
// is bound by the two dependent L2 round trips per token (bucket start, then entries), not by
// the rounds, so the smaller tables stay.
#ifndef LZ77_BIG_KEY0
#define LZ77_BIG_KEY0 7
#endif
#ifndef LZ77_BIG_KEY1
#define LZ77_BIG_KEY1 6
#endif
constexpr int kBigB0Bits = LZ77_BIG_KEY0, kBigB1Bits = LZ77_BIG_KEY1;
constexpr int kBigBuckets = 1 << (kBigB0Bits + kBigB1Bits);

__device__ __forceinline__ int big_key(uint32_t b0, uint32_t b1)
{
    return (int)(((b0 & ((1u << kBigB0Bits) - 1u)) << kBigB1Bits) |
                 (b1 & ((1u << kBigB1Bits) - 1u)));
}
#ifndef LZ77_SORT_THREADS
#define LZ77_SORT_THREADS 1024
#endif
constexpr int kSortThreads = LZ77_SORT_THREADS;
constexpr int kSortWarps = kSortThreads / 32;
#ifndef LZ77_SORT_BALLOT
#define LZ77_SORT_BALLOT 1
#endif

// ---------------------------------------------------------------------------
// per-block stable sort of positions by key
// ---------------------------------------------------------------------------

constexpr int kPosBits = 19;                     // positions inside a block (<= 512 KiB)
constexpr uint32_t kPosMask = (1u << kPosBits) - 1u;

// One stable counting-sort pass over n elements.  Element i is elem_of(i); its
// digit is digit_of(element) and store_of(element) is what lands in dst.  Warp w
// handles the contiguous run [w*chunk, (w+1)*chunk) so equal digits keep order.
template <int kBins, typename ElemFn, typename DigitFn, typename StoreFn>
__device__ __forceinline__ void radix_pass(uint32_t *dst, int n, uint32_t *cnt, uint32_t *bin_start,
                                           uint32_t *s_warp, uint32_t *s_total, ElemFn elem_of,
                                           DigitFn digit_of, StoreFn store_of)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int rows = (n + kSortThreads - 1) / kSortThreads;
    const int cbase = warp * rows * 32;
    const int cend = min(n, cbase + rows * 32);

    for (int i = threadIdx.x; i < kSortWarps * kBins; i += kSortThreads) cnt[i] = 0u;
    __syncthreads();
    for (int i = cbase + lane; i < cend; i += 32)
        atomicAdd(&cnt[warp * kBins + digit_of(elem_of(i))], 1u);
    __syncthreads();
    // exclusive scan down the warps of every bin, then across the bins
    uint32_t tot = 0;
    if (threadIdx.x < kBins) {
        for (int w = 0; w < kSortWarps; w++) {
            const uint32_t v = cnt[w * kBins + threadIdx.x];
            cnt[w * kBins + threadIdx.x] = tot;
            tot += v;
        }
    }
    const uint32_t base = block_exclusive_scan_u32<kSortThreads>(tot, s_warp, s_total);
    if (threadIdx.x < kBins) bin_start[threadIdx.x] = base;
    __syncthreads();
    for (int r = 0; r < rows; r++) {
        const int i = cbase + r * 32 + lane;
        const bool valid = i < cend;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const uint32_t el = elem_of(i);
            const int d = digit_of(el);
#if LZ77_SORT_BALLOT
            // lanes with the same digit, from one ballot per digit bit: MATCH.ANY occupies its
            // unit for ~55 cycles a row, and this kernel has issue slots to spare (17 % busy)
            unsigned peers = vmask;
#pragma unroll
            for (int bit = 0; (1 << bit) < kBins; bit++) {
                const bool one = (d >> bit) & 1;
                const unsigned bal = __ballot_sync(vmask, one);
                peers &= one ? bal : ~bal;
            }
#else
            const unsigned peers = __match_any_sync(vmask, d);
#endif
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) old = atomicAdd(&cnt[warp * kBins + d], (uint32_t)__popc(peers));
            old = __shfl_sync(peers, old, leader);
            dst[bin_start[d] + old + __popc(peers & lt_mask)] = store_of(el);
        }
    }
    __threadfence_block();
    __syncthreads();
}

// The block is read straight from HBM/L2 (sequentially, once per pass 1): pass 1
// orders the positions by the low digit x[q+1] and stores them with the high
// digit x[q] packed above the position, so pass 2 needs no data access.
#ifndef LZ77_SORT_MINBLOCKS
#define LZ77_SORT_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(kSortThreads, LZ77_SORT_MINBLOCKS)
lz77_block_sort_kernel(const uint8_t *__restrict__ in, long long n, int block_shift,
                       uint32_t *__restrict__ sorted, uint32_t *__restrict__ tmp,
                       uint32_t *__restrict__ bstart)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    __shared__ uint32_t bin_start[(1 << kBigB0Bits) + 1];

    const long long block_bytes = 1LL << block_shift;
    const long long blk_lo = (long long)blockIdx.x << block_shift;
    const int nb = (int)min(block_bytes, n - blk_lo);
    const uint8_t *data = in + blk_lo;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem);  // [32 warps][bins of the wider digit]
    uint32_t *my_sorted = sorted + blk_lo;
    uint32_t *my_tmp = tmp + blk_lo;
    uint32_t *my_bstart = bstart + (long long)blockIdx.x * (kBigBuckets + 1);
    // byte q+1 of the last position of the input does not exist: it reads as 0
    auto byte_at = [&](int q) -> uint32_t { return blk_lo + q < n ? (uint32_t)data[q] : 0u; };

    // pass 1: low digit = x[q+1], entries leave as q | x[q] << 19
    radix_pass<1 << kBigB1Bits>(
        my_tmp, nb, cnt, bin_start, s_warp, &s_total,
        [&](int i) { return (uint32_t)i; },
        [&](uint32_t q) { return (int)(byte_at((int)q + 1) & ((1u << kBigB1Bits) - 1u)); },
        [&](uint32_t q) { return q | ((byte_at((int)q) & ((1u << kBigB0Bits) - 1u)) << kPosBits); });
    // pass 2: high digit, carried in the entry
    radix_pass<1 << kBigB0Bits>(
        my_sorted, nb, cnt, bin_start, s_warp, &s_total,
        [&](int i) { return my_tmp[i]; },
        [&](uint32_t el) { return (int)(el >> kPosBits); },
        [&](uint32_t el) { return el & kPosMask; });
    // bucket starts from the sorted list: bucket k starts at the first entry whose key is
    // >= k (entry i writes the starts of the keys in (key[i-1], key[i]]; the last entry also
    // those behind it).  The list was written by this CTA (fence + barrier in radix_pass).
    auto key_of = [&](int i) -> int {
        const int q = (int)my_sorted[i];
        return big_key(byte_at(q), byte_at(q + 1));
    };
    if (nb == 0) {
        for (int k = threadIdx.x; k <= kBigBuckets; k += kSortThreads) my_bstart[k] = 0u;
        return;
    }
    // short runs of empty buckets are written by the entry's thread, long ones (text uses a
    // fraction of the 65536 pairs) are queued and filled by the whole CTA
    constexpr int kGapInline = 32, kGapQueue = kBigBuckets / kGapInline + 2;
    static_assert(kGapQueue * 3 <= kSortWarps * (1 << (kBigB0Bits > kBigB1Bits ? kBigB0Bits : kBigB1Bits)),
                  "the gap queue lives in the counter area");
    int *gap_q = reinterpret_cast<int *>(cnt);
    if (threadIdx.x == 0) s_total = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += kSortThreads) {
        const int k_hi = key_of(i);
        const int k_lo = i > 0 ? key_of(i - 1) + 1 : 0;
        if (k_hi - k_lo < kGapInline) {
            for (int k = k_lo; k <= k_hi; k++) my_bstart[k] = (uint32_t)i;
        } else {
            const int slot = (int)atomicAdd(&s_total, 1u);
            gap_q[3 * slot] = k_lo, gap_q[3 * slot + 1] = k_hi, gap_q[3 * slot + 2] = i;
        }
        if (i == nb - 1) {  // behind the last entry: every remaining bucket is empty
            const int slot = (int)atomicAdd(&s_total, 1u);
            gap_q[3 * slot] = k_hi + 1, gap_q[3 * slot + 1] = kBigBuckets, gap_q[3 * slot + 2] = nb;
        }
    }
    __syncthreads();
    const int n_gaps = (int)s_total;
    for (int g = 0; g < n_gaps; g++) {
        const int k_lo = gap_q[3 * g], k_hi = gap_q[3 * g + 1];
        const uint32_t v = (uint32_t)gap_q[3 * g + 2];
        for (int k = k_lo + (int)threadIdx.x; k <= k_hi; k += kSortThreads) my_bstart[k] = v;
    }
}

// ---------------------------------------------------------------------------
// parse with block-level buckets
// ---------------------------------------------------------------------------

// first index in [0, n) of the ascending list e[] whose value is >= lo (n if none)
__device__ __forceinline__ int warp_lower_bound_g(const uint32_t *__restrict__ e, int n, int lo,
                                                  int lane)
{
    int base = 0, cnt = n;
    while (cnt > 32) {
        const int step = (cnt + 31) >> 5;
        const int idx = base + lane * step;
        const bool ge = idx < base + cnt ? (int)__ldg(e + idx) >= lo : true;
        const unsigned m = __ballot_sync(0xffffffffu, ge);
        const int first = m ? __ffs(m) - 1 : 32;
        if (first == 0) return base;
        const int nb = base + (first - 1) * step + 1;
        const int ne = min(base + cnt, base + first * step + 1);
        base = nb;
        cnt = ne - nb;
    }
    const int idx = base + lane;
    const bool ge = idx < base + cnt ? (int)__ldg(e + idx) >= lo : true;
    const unsigned m = __ballot_sync(0xffffffffu, ge);
    return base + (m ? __ffs(m) - 1 : 32);
}

// One candidate per lane of a round.  The first four bytes are compared without a
// branch by every lane (a lane without a candidate, `in` false, reads the target
// itself and is masked afterwards): most candidates end there, and a divergent round
// costs every path once.  Lanes whose first word matches go on from byte 4.
template <bool kSmallLA>
__device__ __forceinline__ int round_match_len_big(const uint8_t *smem, int q, bool in, int p0,
                                                   const uint32_t (&tgt)[4], int max_len)
{
    const int qq = in ? q : p0;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(smem + (qq & ~3));
    const int sh = (qq & 3) * 8;
    const uint32_t a1 = w[1];
    const uint32_t x = __funnelshift_r(w[0], a1, sh) ^ tgt[0];
    int l = (int)min((uint32_t)(__ffs(x) - 1) >> 3, 4u);  // __ffs(0) - 1 wraps: 4
    if (in && x == 0u) l = match_len_from4<kSmallLA>(smem, w, sh, a1, p0, tgt, max_len);
    return in ? min(l, max_len) : 0;
}

constexpr int kBigWarps = 16;
#ifndef LZ77_BIG_TOKLOOP
#define LZ77_BIG_TOKLOOP 2  // 2: token loop with a packed running best (see search_bucket.cu); 1: round-1 loop
#endif

template <bool kSmallLA>
__global__ void __launch_bounds__(kBigWarps * 32, 2)
lz77_parse_bigwin_kernel(const uint8_t *__restrict__ in, long long n, long long pre, int lead,
                         Params P, int hist_cap, const uint32_t *__restrict__ sorted,
                         const uint32_t *__restrict__ bstart, uint32_t *__restrict__ tok_tmp,
                         uint32_t *__restrict__ seg_ntok)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;

    constexpr int kThreads = kBigWarps * 32;
    constexpr int tile_bytes = kBigWarps * kSegBytes;
    const int lane = threadIdx.x & 31;
    // (through REDUX: known to be uniform, so the segment bounds and the parse position
    // derived from it stay off the vector ALU)
    const int warp = LZ77_BIG_TOKLOOP == 2 ? (int)__reduce_max_sync(0xffffffffu, threadIdx.x >> 5)
                                           : (int)(threadIdx.x >> 5);
    const long long tile_lo = (long long)blockIdx.x * tile_bytes;
    const long long blk_i = tile_lo >> P.block_shift;
    const long long blk_lo = blk_i << P.block_shift;

    // independent blocks: the window starts at the block; history mode: it slides across
    // block seams into the `pre` valid bytes in front of `in` (lz77.c:101-105)
    long long hist = P.history ? tile_lo + pre : tile_lo - blk_lo;
    if (hist > P.window) hist = P.window;
    const int hist_al = (int)((hist + 15) & ~15LL);
    const long long src_lo = tile_lo - hist_al;
    long long src_hi = tile_lo + tile_bytes;
    if (src_hi > n) src_hi = n;
    const int bytes = (int)(src_hi - src_lo);
    const int bulk = bytes & ~15;
    const int dst0 = hist_cap - hist_al;  // smem index of global byte src_lo

    if (threadIdx.x == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && bulk > 0) {
        mbar_expect_tx(&mbar, (uint32_t)bulk);
        tma_load_1d(smem + dst0, in + src_lo, (uint32_t)bulk, &mbar);
    }
    for (int i = bulk + threadIdx.x; i < bytes + 64; i += kThreads)
        smem[dst0 + i] = (i < bytes) ? in[src_lo + i] : (uint8_t)0;
    if (bulk > 0) mbar_wait(&mbar, 0);
    __syncthreads();

    const long long seg_lo = tile_lo + (long long)warp * kSegBytes;
    if (seg_lo >= n) return;
    const long long sgm = seg_lo / kSegBytes;
    long long seg_hi = seg_lo + kSegBytes;
    if (seg_hi > n) seg_hi = n;
    const int seg_end = (int)(seg_hi - src_lo) + dst0;
    int p0 = (int)(seg_lo - src_lo) + dst0;
    const int blk_idx = (int)(blk_lo - src_lo) + dst0;  // smem index of block byte 0 (may be < 0)
    const int first_idx = dst0 + hist_al - (int)hist;   // oldest byte a match may start at
    // bucket tables of this block and (history mode) of the block in front of it; table
    // slot 0 belongs to the `lead` block sorted in front of the piece
    const long long slot = blk_i + lead;
    const uint32_t *blk_sorted = sorted + (slot << P.block_shift);
    const uint32_t *blk_bstart = bstart + slot * (kBigBuckets + 1);
    const bool has_prev = P.history && slot >= 1 && first_idx < blk_idx;
    const uint32_t *prev_sorted = blk_sorted - (1LL << P.block_shift);
    const uint32_t *prev_bstart = blk_bstart - (kBigBuckets + 1);
    const int prev_idx = blk_idx - (int)(1LL << P.block_shift);  // smem index of its byte 0
    uint32_t *const tok_row = tok_tmp + sgm * kSegBytes;
    uint32_t *tok_at = tok_row;  // one 4-byte store per token by one lane (L2 merges the sectors)
    const int len_shift = P.ob, lit_shift = P.ob + P.lb;
    const int la = P.la, window = P.window;


    if (LZ77_BIG_TOKLOOP == 2) {
        // The vector ALU is the busiest unit of the parse kernels (ncu), so the running best of
        // a lane is ONE packed key, len * 2^17 - start: a multiply-add (FMA pipe) and a max per
        // round instead of a compare and two selects; equal lengths prefer the older start for
        // free; the REDUX result unpacks on the uniform datapath together with the loop
        // control; the token is two multiply-adds.  The last byte of a segment (no match
        // possible, tree.c:136) leaves the loop; a position without reach finds no candidate
        // inside its window and falls through as a literal.
        const uint32_t sdata = smem_u32(smem);
        constexpr int kMul = 1 << 17;      // staged indices stay below 2^17 (64 KiB + tile)
        constexpr int kNone = -(kMul - 1);  // length 0, no start
        const int la1 = la - 1;
        const int last = seg_end - 1;
        uint32_t len_mul = 1u << len_shift, lit_mul = 1u << lit_shift;
        asm("" : "+r"(len_mul), "+r"(lit_mul));  // (opaque: keeps the multiplies)
        int ntok = 0;
        while (p0 < last) {
            const int max_len = min(la1, last - p0);         // lz77.c:87,134 + tree.c:136
            const int lo_idx = max(p0 - window, first_idx);  // lz77.c:101-105
            uint32_t tgt[4];
            {
                const uint32_t *w = reinterpret_cast<const uint32_t *>(smem + (p0 & ~3));
                const int sh = p0 * 8;  // (the funnel shift wraps: only bits 3..4 count)
                const uint32_t a0 = w[0], a1 = w[1];
                const uint32_t a2 = w[2], a3 = w[3], a4 = w[4];
                tgt[0] = __funnelshift_r(a0, a1, sh);
                tgt[1] = __funnelshift_r(a1, a2, sh);
                tgt[2] = __funnelshift_r(a2, a3, sh);
                tgt[3] = __funnelshift_r(a3, a4, sh);
            }
            const uint32_t b0 = tgt[0] & 0xffu;
            int best = kNone;
            if (max_len >= 2 && lo_idx < p0) {
                const int key = big_key(b0, tgt[0] >> 8);
                const int full_key = max_len * kMul - (kMul - 1);  // any key of maximum length
                auto scan_list = [&](const uint32_t *list, const uint32_t *starts, int base_idx) {
                    const int bs = (int)__ldg(starts + key);
                    const int bn = (int)__ldg(starts + key + 1) - bs;
                    const uint32_t *e = list + bs;
                    const int lo_blk = lo_idx - base_idx, p_blk = p0 - base_idx;
                    int i = bn <= 64 ? 0 : warp_lower_bound_g(e, bn, lo_blk, lane);
                    for (; i < bn; i += 32) {
                        const int idx = i + lane;
                        const int qb = idx < bn ? (int)__ldg(e + idx) : 0x7fffffff;
                        const bool in = qb >= lo_blk && qb < p_blk;
                        const int q = in ? qb + base_idx : 0;
                        const int l = round_match_len_big<kSmallLA>(smem, q, in, p0, tgt, max_len);
                        best = max(best, in ? l * kMul - q : kNone);
                        if (__any_sync(0xffffffffu, best >= full_key || qb >= p_blk)) break;
                    }
                };
                // oldest first: the tail of the previous block's list, then this block's
                // (not needed once a maximum-length match is in: a later one is not longer)
                bool full = false;
                if (has_prev && lo_idx < blk_idx) {
                    scan_list(prev_sorted, prev_bstart, prev_idx);
                    full = __any_sync(0xffffffffu, best >= full_key);
                }
                if (!full) scan_list(blk_sorted, blk_bstart, blk_idx);
            }
            const int kbest = __reduce_max_sync(0xffffffffu, best);
            int len = (kbest + (kMul - 1)) >> 17;
            int off = p0 + kbest - (len << 17);  // p0 - start
            if (len < 2) {
                // length 1 (such bytes sit in many buckets): scan of the staged window
                const int q1 = oldest_byte_match(sdata, lo_idx, p0, b0, lane);
                len = q1 >= 0 ? 1 : 0;
                off = q1 >= 0 ? p0 - q1 : 0;
            }
            const uint32_t lit = smem[p0 + len];
            const uint32_t tok = lit * lit_mul + ((uint32_t)len * len_mul + (uint32_t)off);
            if (lane == 0) tok_row[ntok] = tok;
            ntok++;
            p0 += len + 1;
        }
        if (p0 == last) {  // the segment's last byte: a literal
            if (lane == 0) tok_row[ntok] = (uint32_t)smem[p0] << lit_shift;
            ntok++;
            p0++;
        }
        tok_at = tok_row + ntok;
    }
    while (LZ77_BIG_TOKLOOP != 2 && p0 < seg_end) {
        const int max_len = min(la, seg_end - p0) - 1;  // lz77.c:87,134 + tree.c:136
        const int reach = min(p0 - first_idx, window);  // lz77.c:101-105
        int len = 0, off = 0;

        if (max_len > 0 && reach > 0) {
            const int lo_idx = p0 - reach;
            uint32_t tgt[4];
            {
                const uint32_t *w = reinterpret_cast<const uint32_t *>(smem + (p0 & ~3));
                const int sh = (p0 & 3) * 8;
                const uint32_t a0 = w[0], a1 = w[1];
                const uint32_t a2 = w[2], a3 = w[3], a4 = w[4];
                tgt[0] = __funnelshift_r(a0, a1, sh);
                tgt[1] = __funnelshift_r(a1, a2, sh);
                tgt[2] = __funnelshift_r(a2, a3, sh);
                tgt[3] = __funnelshift_r(a3, a4, sh);
            }
            const uint32_t b0 = tgt[0] & 0xffu;
            int best_len = 0, best_q = 0;
            if (max_len >= 2) {
                const int key = big_key(b0, tgt[0] >> 8);
                // one bucket list: entries are positions in their block, ascending; smem
                // index = entry + base_idx
                auto scan_list = [&](const uint32_t *list, const uint32_t *starts, int base_idx) {
                    const int bs = (int)__ldg(starts + key);
                    const int bn = (int)__ldg(starts + key + 1) - bs;
                    const uint32_t *e = list + bs;
                    const int lo_blk = lo_idx - base_idx, p_blk = p0 - base_idx;
                    int i = bn <= 64 ? 0 : warp_lower_bound_g(e, bn, lo_blk, lane);
                    for (; i < bn; i += 32) {
                        const int idx = i + lane;
                        const int qb = idx < bn ? (int)__ldg(e + idx) : 0x7fffffff;
                        const bool in = qb >= lo_blk && qb < p_blk;
                        const int q = in ? qb + base_idx : 0;
                        const int l = round_match_len_big<kSmallLA>(smem, q, in, p0, tgt, max_len);
                        if (l > best_len) {  // nearer than anything this lane has seen: longer
                            best_len = l;
                            best_q = q;
                        }
                        if (__any_sync(0xffffffffu, best_len >= max_len || qb >= p_blk)) break;
                    }
                };
                // oldest first: the tail of the previous block's list, then this block's
                // (not needed once a maximum-length match is in: a later one is not longer)
                bool full = false;
                if (has_prev && lo_idx < blk_idx) {
                    scan_list(prev_sorted, prev_bstart, prev_idx);
                    full = __any_sync(0xffffffffu, best_len >= max_len);
                }
                if (!full) scan_list(blk_sorted, blk_bstart, blk_idx);
            }
            // (no candidate: length 0 in the top bits, the start is not used)
            const uint32_t k = __reduce_max_sync(
                0xffffffffu, ((uint32_t)best_len << 20) | (0xfffffu - (uint32_t)best_q));
            len = (int)(k >> 20);
            int q_best = (int)(0xfffffu - (k & 0xfffffu));
            if (len < 2) {
                // length 1: the oldest byte of the window equal to the first lookahead
                // byte -- forward SWAR scan of the staged window, 512 bytes per step
                const uint32_t b0x4 = b0 * 0x01010101u;
                int q1 = 0x7fffffff;
                for (int base = lo_idx & ~15; base < p0; base += 512) {
                    const int g = base + lane * 16;
                    if (g < p0) {
                        const uint4 v = *reinterpret_cast<const uint4 *>(smem + g);
                        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int i = 3; i >= 0; i--) {
                            uint32_t m = zero_bytes(wv[i] ^ b0x4);
                            while (m) {
                                const int bit = __ffs(m) - 1;
                                m ^= 1u << bit;
                                const int q = g + 4 * i + (bit >> 3);
                                if (q >= lo_idx && q < p0 && q < q1) q1 = q;
                            }
                        }
                    }
                    if (__any_sync(0xffffffffu, q1 != 0x7fffffff)) break;
                }
                q1 = (int)__reduce_min_sync(0xffffffffu, (unsigned)q1);
                len = q1 != 0x7fffffff ? 1 : 0;
                q_best = q1;
            }
            off = len ? p0 - q_best : 0;
        }

        const uint32_t lit = smem[p0 + len];
        const uint32_t tok = (uint32_t)off | ((uint32_t)len << len_shift) | (lit << lit_shift);
        if (lane == 0) *tok_at = tok;
        tok_at++;
        p0 += len + 1;
    }
    if (lane == 0) seg_ntok[sgm] = (uint32_t)(tok_at - tok_row);
}

// ---------------------------------------------------------------------------

#ifndef LZ77_BIG_PIECE_MIB
#define LZ77_BIG_PIECE_MIB 64
#endif
constexpr long long kBigPiece = (long long)LZ77_BIG_PIECE_MIB << 20;  // bucket tables are built for this much input at a time

static size_t bigwin_piece_scratch(long long n_in, const Params &P)
{
    // (history mode sorts one more block in front of every piece)
    const long long n_blocks = ((n_in + P.block - 1) >> P.block_shift) + (P.history ? 1 : 0);
    const long long npos = n_blocks << P.block_shift;
    return ((size_t)npos * 4 * 2 + (size_t)n_blocks * (kBigBuckets + 1) * 4 + 4095) & ~(size_t)4095;
}

// two table sets: piece k+1 is sorted while piece k is parsed
size_t bigwin_scratch_bytes(long long n_in, const Params &P)
{
    return 2 * bigwin_piece_scratch(n_in < kBigPiece ? n_in : kBigPiece, P) + 4096;
}

namespace {
struct BigwinStreams {
    bool ready = false;
    cudaStream_t sort = nullptr;   // high priority: its few CTAs slip in between parse CTAs
    cudaEvent_t fork = nullptr, sorted[2] = {nullptr, nullptr}, parsed[2] = {nullptr, nullptr};
};
BigwinStreams g_bw_dev[16];  // per device (one host thread drives one device)

cudaError_t bigwin_streams(BigwinStreams **out)
{
    int dev = 0;
    cudaError_t rc = cudaGetDevice(&dev);
    if (rc != cudaSuccess) return rc;
    if (dev < 0 || dev >= 16) return cudaErrorInvalidDevice;
    BigwinStreams &bw = g_bw_dev[dev];
    *out = &bw;
    if (bw.ready) return cudaSuccess;
    int lo_pri = 0, hi_pri = 0;
    rc = cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
    if (rc != cudaSuccess) return rc;
    rc = cudaStreamCreateWithPriority(&bw.sort, cudaStreamNonBlocking, hi_pri);
    if (rc != cudaSuccess) return rc;
    cudaEventCreateWithFlags(&bw.fork, cudaEventDisableTiming);
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&bw.sorted[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&bw.parsed[i], cudaEventDisableTiming);
    }
    bw.ready = true;
    return cudaGetLastError();
}
}  // namespace

// called by lz77_gpu_shutdown() with the device current
void bigwin_release(int device)
{
    if (device < 0 || device >= 16 || !g_bw_dev[device].ready) return;
    BigwinStreams &bw = g_bw_dev[device];
    cudaStreamDestroy(bw.sort);
    cudaEventDestroy(bw.fork);
    for (int i = 0; i < 2; i++) {
        cudaEventDestroy(bw.sorted[i]);
        cudaEventDestroy(bw.parsed[i]);
    }
    bw = BigwinStreams();
}

// d_in points at a block boundary.  The input is handled in pieces of kBigPiece
// bytes: the block sort of piece k+1 runs on a high-priority side stream while
// piece k is parsed on `st` (two sets of tables in the scratch area).
cudaError_t launch_parse_bigwin(const uint8_t *d_in, long long n_in, long long pre,
                                const Params &P, void *scratch, uint32_t *tok_tmp,
                                uint32_t *seg_ntok, cudaStream_t st)
{
    if (n_in <= 0) return cudaSuccess;
    BigwinStreams *bwp = nullptr;
    cudaError_t rc = bigwin_streams(&bwp);
    if (rc != cudaSuccess) return rc;
    BigwinStreams &g_bw = *bwp;
    const size_t piece_scratch = bigwin_piece_scratch(n_in < kBigPiece ? n_in : kBigPiece, P);
    constexpr int kWideBins = 1 << (kBigB0Bits > kBigB1Bits ? kBigB0Bits : kBigB1Bits);
    const size_t sort_smem = (size_t)kSortWarps * kWideBins * 4;
    rc = cudaFuncSetAttribute(lz77_block_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)sort_smem);
    if (rc != cudaSuccess) return rc;
    const int hist_cap = (P.window + 15) & ~15;
    const long long tile_bytes = (long long)kBigWarps * kSegBytes;
    const size_t parse_smem = (size_t)hist_cap + (size_t)tile_bytes + 128;
    auto parse = P.la <= 16 ? lz77_parse_bigwin_kernel<true> : lz77_parse_bigwin_kernel<false>;
    rc = cudaFuncSetAttribute(parse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)parse_smem);
    if (rc != cudaSuccess) return rc;

    cudaEventRecord(g_bw.fork, st);
    cudaStreamWaitEvent(g_bw.sort, g_bw.fork, 0);
    long long k = 0;
    for (long long o = 0; o < n_in; o += kBigPiece, k++) {
        const long long len = n_in - o < kBigPiece ? n_in - o : kBigPiece;
        const int buf = (int)(k & 1);
        // history mode: the block in front of the piece is sorted along with it, so the
        // first block of the piece finds the candidates of its window there
        const int lead = (P.history && pre + o >= P.block) ? 1 : 0;
        const long long n_blocks = ((len + P.block - 1) >> P.block_shift) + lead;
        const long long npos = n_blocks << P.block_shift;
        uint32_t *sorted = (uint32_t *)((char *)scratch + buf * piece_scratch);
        uint32_t *tmp = sorted + npos;
        uint32_t *bstart = tmp + npos;
        if (k >= 2) cudaStreamWaitEvent(g_bw.sort, g_bw.parsed[buf], 0);  // tables free again
        lz77_block_sort_kernel<<<(unsigned)n_blocks, kSortThreads, sort_smem, g_bw.sort>>>(
            d_in + o - lead * P.block, len + lead * P.block, P.block_shift, sorted, tmp, bstart);
        cudaEventRecord(g_bw.sorted[buf], g_bw.sort);
        cudaStreamWaitEvent(st, g_bw.sorted[buf], 0);
        const long long n_tiles = (len + tile_bytes - 1) / tile_bytes;
        parse<<<(unsigned)n_tiles, kBigWarps * 32, parse_smem, st>>>(
            d_in + o, len, pre + o, lead, P, hist_cap, sorted, bstart, tok_tmp + o,
            seg_ntok + o / kSegBytes);
        cudaEventRecord(g_bw.parsed[buf], st);
    }
    return cudaGetLastError();
}

}  // namespace lz77
