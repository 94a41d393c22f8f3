// encode.cu -- block-parallel LZ77 encoder kernels for sm_100a.
//
// Replaces, for a batch of independent blocks at once:
//   tree.c:62-260 (BST insert/find/delete/updateOffset)  -> shared-memory
//       windowed warp search over position buckets: search_bucket.cu (windows
//       <= 8191) and search_bigwin.cu (larger windows)
//   lz77.c:89-135 (greedy token loop, match() lz77.c:209-224) -> one warp per
//       parse segment, sequential in the segment, parallel across segments
//   lz77.c:246-252 writecode + bitio.c:203-239 bitIO_write -> warp-cooperative
//       bit-packer (lz77_pack_kernel)
//
// The result is the stream oracle/lz77_oracle.c:lz77o_segmented_encode() defines,
// byte for byte: inside a block the longest match (<= min(LA, bytes left in the
// segment) - 1) against the last min(window, position in block) bytes, farthest
// offset among the longest (keeps the decoder's dependency chains short), then
// a literal.
#include "kernels.cuh"

namespace lz77 {

// ---------------------------------------------------------------------------
// token-count prefix sums (uint32 counts -> uint64 exclusive prefix)
// ---------------------------------------------------------------------------

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v,
                                                                   unsigned long long *total)
{
    __shared__ unsigned long long warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        unsigned long long s = lane < nw ? warp_sums[lane] : 0ull;
        unsigned long long sinc = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, sinc, d);
            if (lane >= d) sinc += t;
        }
        warp_sums[lane] = sinc - s;  // exclusive warp offsets
        if (lane == 31) *total = sinc;
    }
    __syncthreads();
    unsigned long long r = warp_sums[warp] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads)
scan_partial_kernel(const uint32_t *__restrict__ in, long long n,
                    unsigned long long *__restrict__ partial)
{
    __shared__ unsigned long long total;
    const long long base = (long long)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++)
        if (base + i < n) s += in[base + i];
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
scan_top_kernel(unsigned long long *__restrict__ partial, long long n_partial,
                unsigned long long *__restrict__ grand_total,
                unsigned long long *host_total /* mapped pinned memory, may be null */)
{
    __shared__ unsigned long long total;
    // *grand_total carries the token count of the chunks encoded before this one
    unsigned long long carry = *grand_total;
    __syncthreads();
    for (long long base = 0; base < n_partial; base += blockDim.x) {
        const long long i = base + threadIdx.x;
        const unsigned long long v = i < n_partial ? partial[i] : 0ull;
        const unsigned long long ex = block_exclusive_scan(v, &total);
        if (i < n_partial) partial[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *grand_total = carry;
        // the chunked host path reads the running count straight from pinned memory: a
        // D2H memcpy of 8 bytes would queue behind the bulk output copies
        if (host_total) {
            *reinterpret_cast<volatile unsigned long long *>(host_total) = carry;
            __threadfence_system();
        }
    }
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const uint32_t *__restrict__ in, long long n,
                  const unsigned long long *__restrict__ partial,
                  unsigned long long *__restrict__ out)
{
    __shared__ unsigned long long total;
    const long long base = (long long)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = base + i < n ? in[base + i] : 0u;
        s += v[i];
    }
    unsigned long long run = partial[blockIdx.x] + block_exclusive_scan(s, &total);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

// ---------------------------------------------------------------------------
// K2: warp-cooperative bit-packer
// ---------------------------------------------------------------------------
//
// Segment s owns stream bits [32 + T*prefix[s], 32 + T*(prefix[s] + ntok[s])).
// 32-bit words that lie fully inside that range are assembled in registers
// from the <= 5 tokens that overlap them and stored (four at a time as one
// 128-bit store when the lane's group is interior and 16-byte aligned); the
// first and last word of a segment may be shared with its neighbours and are
// OR-ed into words lz77_pack_prepare_kernel zeroed.

template <int kT>  // token width when known at compile time (24, 32), 0 = any
__device__ __forceinline__ uint32_t gather_word(const uint32_t *__restrict__ toks, int n_tok,
                                                int rel_lo, int tbits_rt)
{
    const int tbits = kT ? kT : tbits_rt;  // divisions by a constant when kT != 0
    // rel_lo: bit offset of the word relative to the segment's first token bit
    int t0 = rel_lo <= 0 ? 0 : rel_lo / tbits;
    int t1 = (rel_lo + 31) / tbits;
    if (t1 > n_tok - 1) t1 = n_tok - 1;
    uint32_t val = 0;
    for (int t = t0; t <= t1; t++) {
        const int tb = t * tbits - rel_lo;  // in (-tbits, 32)
        const uint32_t v = __ldg(toks + t);
        val |= tb >= 0 ? v << tb : v >> (-tb);
    }
    return val;
}

// Zeroes the LAST word of every segment.  A segment's first word is either
// word-aligned (then it is an interior or last word) or it is the last word of
// the segment before it -- possibly one encoded by an earlier chunk, which has
// already OR-ed its bits in, so it must not be zeroed again.
__global__ void lz77_pack_prepare_kernel(const uint32_t *__restrict__ seg_ntok,
                                         const unsigned long long *__restrict__ prefix,
                                         long long n_seg, Params P, bool write_header,
                                         uint32_t *__restrict__ out)
{
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0 && write_header)
        out[0] = (uint32_t)P.sb | ((uint32_t)P.la << 16);  // lz77.c:74-75
    if (s >= n_seg) return;
    const uint32_t nt = seg_ntok[s];
    if (!nt) return;
    const long long b0 = kHeaderBits + (long long)P.tbits * (long long)prefix[s];
    const long long b1 = b0 + (long long)P.tbits * nt;
    // (a first segment of a later chunk that ends inside the word it shares with
    // the previous chunk leaves that word alone)
    const bool shared_with_earlier_chunk =
        s == 0 && !write_header && (b0 & 31) != 0 && ((b1 - 1) >> 5) == (b0 >> 5);
    if (!shared_with_earlier_chunk) out[(b1 - 1) >> 5] = 0;
}

template <int kT>
__global__ void __launch_bounds__(256)
lz77_pack_kernel(const uint32_t *__restrict__ tok_tmp, const uint32_t *__restrict__ seg_ntok,
                 const unsigned long long *__restrict__ prefix, long long n_seg, Params P,
                 uint32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long long s = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n_seg) return;
    const int nt = (int)seg_ntok[s];
    if (!nt) return;
    const uint32_t *toks = tok_tmp + s * kSegBytes;
    const int T = kT ? kT : P.tbits;
    const long long b0 = kHeaderBits + (long long)T * (long long)prefix[s];
    const long long b1 = b0 + (long long)T * nt;
    const long long w_first = b0 >> 5, w_last = (b1 - 1) >> 5;

    // groups of four words, aligned to 16 bytes
    for (long long w4 = (w_first & ~3LL) + 4LL * lane; w4 <= w_last; w4 += 128) {
        const bool interior = (w4 << 5) >= b0 && ((w4 + 4) << 5) <= b1;
        if (interior) {
            uint4 v;
            v.x = gather_word<kT>(toks, nt, (int)(((w4 + 0) << 5) - b0), T);
            v.y = gather_word<kT>(toks, nt, (int)(((w4 + 1) << 5) - b0), T);
            v.z = gather_word<kT>(toks, nt, (int)(((w4 + 2) << 5) - b0), T);
            v.w = gather_word<kT>(toks, nt, (int)(((w4 + 3) << 5) - b0), T);
            *reinterpret_cast<uint4 *>(out + w4) = v;  // coalesced 128-bit store
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const long long w = w4 + i;
                if (w < w_first || w > w_last) continue;
                const uint32_t val = gather_word<kT>(toks, nt, (int)((w << 5) - b0), T);
                if ((w << 5) >= b0 && ((w + 1) << 5) <= b1)
                    out[w] = val;
                else
                    atomicOr(out + w, val);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------


size_t encode_scratch_bytes(long long n_in, const Params &P)
{
    if (parse_bucket_fused(P))  // no unpacked tokens in HBM: look-back words + a spill area
        return 4096 + parse_bucket_fused_scratch(n_in);
    const long long n_seg = (n_in + kSegBytes - 1) / kSegBytes;
    const long long n_part = (n_seg + kScanTile - 1) / kScanTile;
    size_t b = 0;
    b += (size_t)(n_seg * kSegBytes) * sizeof(uint32_t);        // tok_tmp
    b += (size_t)((n_seg + 63) & ~63LL) * sizeof(uint32_t);       // seg_ntok
    b += (size_t)((n_seg + 63) & ~63LL) * sizeof(unsigned long long);  // prefix
    b += (size_t)((n_part + 63) & ~63LL) * sizeof(unsigned long long); // partials
    b += 256;                                                      // grand total
    if (P.window > 8191)                                           // block-level bucket tables
        b += bigwin_scratch_bytes(n_in, P);
    return b + 4096;
}

static inline char *carve(char *&p, size_t bytes)
{
    char *r = p;
    p += (bytes + 255) & ~(size_t)255;
    return r;
}

EncodePlan encode_plan(void *scratch, long long n_in_total, const Params &P)
{
    const long long n_seg = (n_in_total + kSegBytes - 1) / kSegBytes;
    const long long n_part = (n_seg + kScanTile - 1) / kScanTile;
    char *p = (char *)scratch;
    EncodePlan pl;
    pl.n_total = n_in_total;
    pl.fused = nullptr;
    if (parse_bucket_fused(P)) {
        pl.tok_tmp = pl.seg_ntok = nullptr;
        pl.prefix = pl.partial = nullptr;
        pl.total = (unsigned long long *)p;
        pl.fused = p + 256;
        pl.big = nullptr;
        return pl;
    }
    pl.tok_tmp = (uint32_t *)carve(p, (size_t)(n_seg * kSegBytes) * sizeof(uint32_t));
    pl.seg_ntok = (uint32_t *)carve(p, (size_t)n_seg * sizeof(uint32_t));
    pl.prefix = (unsigned long long *)carve(p, (size_t)n_seg * sizeof(unsigned long long));
    pl.partial = (unsigned long long *)carve(p, (size_t)n_part * sizeof(unsigned long long));
    pl.total = (unsigned long long *)carve(p, 8);
    pl.big = P.window > 8191 ? (void *)p : nullptr;
    return pl;
}

// chunks of a chunked encode must start on a block boundary (the largest block size)
long long encode_chunk_granule() { return 524288; }

// Encodes input bytes [lo, lo + n_chunk) (lo a multiple of encode_chunk_granule(),
// hence of the block size; the scans of successive chunks run in stream order, so a
// chunk may reuse the last partial-sum slot of its predecessor) of a buffer whose earlier chunks were encoded by
// earlier calls: tokens are appended behind the *pl.total tokens written so far.
// phase 0: everything; phase 1: the search/parse kernel only; phase 2: the token
// count scan and the bit-packer only (the host pipeline runs the searches of
// consecutive chunks on alternating streams so their tails overlap).
cudaError_t launch_encode_chunk(const uint8_t *d_in_base, long long pre_base, long long lo,
                                long long n_chunk, bool first, const Params &P,
                                const EncodePlan &pl,
                                uint32_t *d_out_words, cudaStream_t st, StageEvents *ev,
                                int phase, unsigned long long *host_total, int slot)
{
    const uint8_t *d_in = d_in_base + lo;
    const long long pre = P.history ? pre_base + lo : 0;  // the earlier chunks are history
    if (pl.fused) {
        // 24-bit tokens, small window: search, parse and pack are one kernel (phase 1)
        if (phase == 2) return cudaSuccess;
        if (ev) cudaEventRecord(ev->e[0], st);
        // (phase 0 = a call of one launch: it resets the look-back state itself; the chunked
        // host path resets once, ordered in front of all its streams -- capi.cu)
        cudaError_t rc = launch_parse_bucket_fused(d_in_base, lo, n_chunk, pl.n_total, pre, first,
                                                   first && phase == 0, slot, P, pl.fused,
                                                   (uint8_t *)d_out_words,
                                                   pl.total, host_total, st);
        if (ev) {
            cudaEventRecord(ev->e[1], st);
            cudaEventRecord(ev->e[2], st);
            cudaEventRecord(ev->e[3], st);
        }
        return rc;
    }
    const long long seg0 = lo / kSegBytes;
    const long long n_seg = (n_chunk + kSegBytes - 1) / kSegBytes;
    const long long n_part = (n_seg + kScanTile - 1) / kScanTile;
    uint32_t *tok_tmp = pl.tok_tmp + seg0 * kSegBytes;
    uint32_t *seg_ntok = pl.seg_ntok + seg0;
    unsigned long long *prefix = pl.prefix + seg0;
    unsigned long long *partial = pl.partial + seg0 / kScanTile;
    unsigned long long *total = pl.total;

    if (first && phase != 1) cudaMemsetAsync(total, 0, 8, st);
    if (ev) cudaEventRecord(ev->e[0], st);
    if (phase == 2) {
        // searched by an earlier phase-1 call
    } else if (n_chunk > 0 && P.window <= 8191) {
        // small windows: bucketed search (search_bucket.cu)
        cudaError_t rc = launch_parse_bucket(d_in, n_chunk, pre, P, tok_tmp, seg_ntok, st);
        if (rc != cudaSuccess) return rc;
    } else if (n_chunk > 0) {
        // large windows: block-level buckets (search_bigwin.cu)
        cudaError_t rc = launch_parse_bigwin(d_in, n_chunk, pre, P, pl.big, tok_tmp, seg_ntok, st);
        if (rc != cudaSuccess) return rc;
    }
    if (ev) cudaEventRecord(ev->e[1], st);
    if (phase == 1) return cudaGetLastError();
    if (n_seg > 0) {
        scan_partial_kernel<<<(unsigned)n_part, kScanThreads, 0, st>>>(seg_ntok, n_seg, partial);
        scan_top_kernel<<<1, 1024, 0, st>>>(partial, n_part, total, host_total);
        scan_apply_kernel<<<(unsigned)n_part, kScanThreads, 0, st>>>(seg_ntok, n_seg, partial,
                                                                      prefix);
    }
    if (ev) cudaEventRecord(ev->e[2], st);
    {
        const long long nthreads = n_seg > 0 ? n_seg : 1;
        lz77_pack_prepare_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(
            seg_ntok, prefix, n_seg, P, first, d_out_words);
        if (n_seg > 0) {
            auto pack = P.tbits == 24 ? lz77_pack_kernel<24>
                                      : P.tbits == 32 ? lz77_pack_kernel<32> : lz77_pack_kernel<0>;
            pack<<<(unsigned)((n_seg + 7) / 8), 256, 0, st>>>(tok_tmp, seg_ntok, prefix, n_seg, P,
                                                               d_out_words);
        }
    }
    if (ev) cudaEventRecord(ev->e[3], st);
    return cudaGetLastError();
}

cudaError_t launch_encode(const uint8_t *d_in, long long n_in, long long pre, const Params &P,
                          void *scratch, uint32_t *d_out_words,
                          unsigned long long **d_total_tokens, cudaStream_t st, StageEvents *ev)
{
    const EncodePlan pl = encode_plan(scratch, n_in, P);
    *d_total_tokens = pl.total;
    return launch_encode_chunk(d_in, pre, 0, n_in, true, P, pl, d_out_words, st, ev, 0, nullptr, 0);
}

int encode_launch_count(long long n_in, const Params &P)
{
    if (n_in <= 0) return 1;
    if (parse_bucket_fused(P)) return 1;  // search + parse + pack in one kernel
    return 6;  // parse, 3 x scan, pack-prepare, pack
}

}  // namespace lz77
