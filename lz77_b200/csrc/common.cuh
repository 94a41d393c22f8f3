// common.cuh -- shared definitions of the sm_100a LZ77 kernels.
//
// Wire format (reference lz77.c:74-75, 246-252; bitio.c:203-239): the stream
// is one little-endian bit string; token k sits at bit 32 + k*T and holds
//   off : OB = bitof(SB) bits | len : LB = bitof(LA) bits | next : 8 bits
// LSB first.  OB <= 16 and LB <= 8, so a token always fits one 32-bit value
//   tok = off | len << OB | next << (OB + LB).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace lz77 {

#ifndef LZ77_SEG_BYTES
#define LZ77_SEG_BYTES 1024
#endif
constexpr int kSegBytes = LZ77_SEG_BYTES;  // greedy parse restarts every segment
constexpr int kHeaderBits = 32;        // SB:16, LA:16 (lz77.c:74-75)

struct Params {
    int sb, la;        // header values
    int ob, lb, tbits; // field widths, token width
    int window;        // usable search reach: min(sb, 2^ob - 1)  (Appendix B2)
    long long block;   // independent block size in bytes (power of two)
    int block_shift;
    int tile_shift;    // decode tile = min(block, 128 KiB): what fits shared memory (a
                       // 512 KiB block is decoded as four tiles)
    int fused_pack;    // encoder: 1 = 24-bit tokens are packed inside the search kernel
                       // (lz77_gpu_set_fused_pack); 0 = 32-bit tokens to scratch + pack kernel
    int history;       // encoder: 1 = the match window slides across block seams like the
                       // reference's (lz77.c:101-105); 0 = independent blocks
};

__host__ __device__ inline int bitof(int n)  // bitio.c:41-43 in integers
{
    int b = 0;
    if (n <= 1) return 0;
    while ((1L << b) < (long)n) b++;
    return b;
}

// ---- unaligned little-endian helpers ------------------------------------

// 32 bits starting at bit `bit` of a word array (guarded at the end)
__device__ __forceinline__ uint32_t load_bits32(const uint32_t *__restrict__ words,
                                                long long n_words, long long bit)
{
    long long w = bit >> 5;
    int s = (int)(bit & 31);
    uint32_t lo = (w < n_words) ? __ldg(words + w) : 0u;
    uint32_t hi = (s != 0 && w + 1 < n_words) ? __ldg(words + w + 1) : 0u;
    return __funnelshift_r(lo, hi, s);
}

// 4 bytes at an arbitrary shared-memory byte index (reads two aligned words)
__device__ __forceinline__ uint32_t lds_u32_unaligned(const uint8_t *smem, int idx)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(smem + (idx & ~3));
    return __funnelshift_r(w[0], w[1], (idx & 3) * 8);
}

// exact per-byte zero detector: bit 7 of every byte of the result is set iff
// that byte of x is zero
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x)
{
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}

// Oldest position q in [lo_idx, p0) of the staged bytes (shared address sdata) whose byte
// equals b0 = the byte at p0, or -1: the length-1 match of the reference (tree.c:118-152 finds
// such a byte when nothing longer exists; ties go to the oldest).  Forward
// SWAR scan by a whole warp, 512 bytes per step; the result is warp-uniform.  The byte at
// p0 itself is the scan's sentinel (it always matches and nothing behind it is looked at),
// so only the first 16-byte chunk needs a bounds mask.
__device__ __forceinline__ int oldest_byte_match(uint32_t sdata, int lo_idx, int p0, uint32_t b0,
                                                 int lane)
{
    const uint32_t b4 = b0 * 0x01010101u;
    for (int base = lo_idx & ~15; base < p0; base += 512) {
        const int g = base + lane * 16;
        uint32_t z0 = 0u, z1 = 0u, z2 = 0u, z3 = 0u;
        if (g <= p0) {
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                         : "r"(sdata + (uint32_t)g));
            z0 = zero_bytes(w0 ^ b4), z1 = zero_bytes(w1 ^ b4);
            z2 = zero_bytes(w2 ^ b4), z3 = zero_bytes(w3 ^ b4);
            if (g < lo_idx) {  // (one lane of the first step) bytes in front of the window
                const int k = lo_idx - g;  // 1..15 of them
                z0 &= k >= 4 ? 0u : 0xffffffffu << (8 * k);
                z1 &= k >= 8 ? 0u : (k > 4 ? 0xffffffffu << (8 * (k - 4)) : 0xffffffffu);
                z2 &= k >= 12 ? 0u : (k > 8 ? 0xffffffffu << (8 * (k - 8)) : 0xffffffffu);
                z3 &= k > 12 ? 0xffffffffu << (8 * (k - 12)) : 0xffffffffu;
            }
        }
        const uint32_t any = z0 | z1 | z2 | z3;
        if (__any_sync(0xffffffffu, any != 0u)) {
            unsigned q = 0x7fffffffu;
            if (any) {
                const uint32_t zs = z0 ? z0 : z1 ? z1 : z2 ? z2 : z3;
                const int wo = z0 ? 0 : z1 ? 4 : z2 ? 8 : 12;
                q = (unsigned)(g + wo + ((__ffs(zs) - 1) >> 3));
            }
            q = __reduce_min_sync(0xffffffffu, q);
            return (int)q < p0 ? (int)q : -1;
        }
    }
    return -1;
}

// ---- mbarrier / TMA bulk copy (cp.async.bulk, SASS UBLKCP) ---------------

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void tma_store_1d(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void tma_store_commit_wait()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// make generic-proxy shared-memory writes visible to the async (TMA) proxy
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// CTA-wide exclusive prefix sum of one value per thread (all threads must call)
template <int kThreads>
__device__ __forceinline__ uint32_t block_exclusive_scan_u32(uint32_t v, uint32_t *s_warp,
                                                             uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = lane < kThreads / 32 ? s_warp[lane] : 0u;
        uint32_t sinc = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, sinc, d);
            if (lane >= d) sinc += t;
        }
        s_warp[lane] = sinc - s;
        if (lane == 31) *total = sinc;
    }
    __syncthreads();
    const uint32_t r = s_warp[warp] + inc - v;
    __syncthreads();
    return r;
}

}  // namespace lz77
