// match.cuh -- common-prefix length of a window candidate and the lookahead,
// the byte loop of the reference's find() (tree.c:136) done a word at a time.
#pragma once

#include "common.cuh"

namespace lz77 {

// Common-prefix length of smem[q..] and smem[p0..] from byte 4 on, capped at max_len,
// for a caller that has already found bytes 0..3 equal.  The first 16 lookahead bytes are
// held in registers as tgt[0..3]; with kSmallLA (LA <= 16) that is the whole lookahead.
// w = the aligned word holding byte q, sh = (q & 3) * 8, a1 = w[1].
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
