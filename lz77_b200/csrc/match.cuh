// match.cuh -- common-prefix length of a window candidate and the lookahead,
// the byte loop of the reference's find() (tree.c:136) done a word at a time.
#pragma once

#include "common.cuh"

namespace lz77 {

// Length of the common prefix of smem[q..] and smem[p0..], capped at max_len.
// The first 16 lookahead bytes are held in registers as tgt[0..3]; with kSmallLA
// (LA <= 16) that is the whole lookahead.
template <bool kSmallLA>
__device__ __forceinline__ int match_len(const uint8_t *smem, int q, int p0,
                                         const uint32_t (&tgt)[4], int max_len)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(smem + (q & ~3));
    const int sh = (q & 3) * 8;
    if (kSmallLA) {
        uint32_t a0 = w[0], a1 = w[1];
        uint32_t x = __funnelshift_r(a0, a1, sh) ^ tgt[0];
        int l;
        if (x) {
            l = (__ffs(x) - 1) >> 3;
        } else {
            uint32_t a2 = w[2];
            x = __funnelshift_r(a1, a2, sh) ^ tgt[1];
            if (x) {
                l = 4 + ((__ffs(x) - 1) >> 3);
            } else {
                uint32_t a3 = w[3];
                x = __funnelshift_r(a2, a3, sh) ^ tgt[2];
                if (x) {
                    l = 8 + ((__ffs(x) - 1) >> 3);
                } else {
                    uint32_t a4 = w[4];
                    x = __funnelshift_r(a3, a4, sh) ^ tgt[3];
                    l = x ? 12 + ((__ffs(x) - 1) >> 3) : 16;
                }
            }
        }
        return min(l, max_len);
    } else {
        // LA > 16: the first 16 bytes against the registers (most candidates differ
        // there), the rest word by word against shared memory
        uint32_t a0 = w[0], a1 = w[1];
        uint32_t x = __funnelshift_r(a0, a1, sh) ^ tgt[0];
        if (x) return min((__ffs(x) - 1) >> 3, max_len);
        uint32_t a2 = w[2];
        x = __funnelshift_r(a1, a2, sh) ^ tgt[1];
        if (x) return min(4 + ((__ffs(x) - 1) >> 3), max_len);
        uint32_t a3 = w[3];
        x = __funnelshift_r(a2, a3, sh) ^ tgt[2];
        if (x) return min(8 + ((__ffs(x) - 1) >> 3), max_len);
        uint32_t a = w[4];
        x = __funnelshift_r(a3, a, sh) ^ tgt[3];
        if (x) return min(12 + ((__ffs(x) - 1) >> 3), max_len);
        int l = 16;
        int wi = 5;
        while (l < max_len) {
            uint32_t b = w[wi++];
            x = __funnelshift_r(a, b, sh) ^ lds_u32_unaligned(smem, p0 + l);
            if (x) {
                l += (__ffs(x) - 1) >> 3;
                break;
            }
            l += 4;
            a = b;
        }
        return min(l, max_len);
    }
}

// The same from byte 4 on, for a caller that has already found bytes 0..3 equal
// (w, sh as above, a1 = w[1]).
template <bool kSmallLA>
__device__ __forceinline__ int match_len_from4(const uint8_t *smem, const uint32_t *w, int sh,
                                               uint32_t a1, int p0, const uint32_t (&tgt)[4],
                                               int max_len)
{
    uint32_t a2 = w[2];
    uint32_t x = __funnelshift_r(a1, a2, sh) ^ tgt[1];
    if (x) return min(4 + ((__ffs(x) - 1) >> 3), max_len);
    uint32_t a3 = w[3];
    x = __funnelshift_r(a2, a3, sh) ^ tgt[2];
    if (x) return min(8 + ((__ffs(x) - 1) >> 3), max_len);
    uint32_t a = w[4];
    x = __funnelshift_r(a3, a, sh) ^ tgt[3];
    if (x) return min(12 + ((__ffs(x) - 1) >> 3), max_len);
    int l = 16;
    if (!kSmallLA) {
        int wi = 5;
        while (l < max_len) {
            uint32_t b = w[wi++];
            x = __funnelshift_r(a, b, sh) ^ lds_u32_unaligned(smem, p0 + l);
            if (x) {
                l += (__ffs(x) - 1) >> 3;
                break;
            }
            l += 4;
            a = b;
        }
    }
    return min(l, max_len);
}

}  // namespace lz77
