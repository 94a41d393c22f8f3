/*
 * TEST INFRASTRUCTURE -- a CPU stand-in for liblz77b200.so, built from the oracle
 * (oracle/lz77_oracle.c), so that the host logic of the command-line program
 * (lz77_b200/csrc/cli/codec.c: pieces, the retained output tail re-encoded as literal
 * tokens, the read-ahead / writer threads, the retry on LZ77_E_SPACE) can be exercised
 * without a GPU.  tests/test_host.py links the unmodified CLI sources against this file in
 * a temporary directory; nothing in the product links, loads or ships it.
 *
 * Only the entry points codec.c calls are provided.  lz77_gpu_encode follows the encoder's
 * specification (lz77o_segmented_encode: blocks + 1 KiB segments), lz77_gpu_decode the
 * reference decoder restatement -- including "LZ77_E_SPACE with the needed size in *n_out".
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "lz77_b200.h"
#include "lz77_oracle.h"

static int bitof_(int n)
{
    int b = 0;
    if (n <= 1)
        return 0;
    while ((1L << b) < (long)n)
        b++;
    return b;
}

int lz77_bitof(int n) { return bitof_(n); }

int lz77_token_bits(int sb, int la)
{
    if (sb == -1) sb = LZ77_DEFAULT_SB;
    if (la == -1) la = LZ77_DEFAULT_LA;
    return bitof_(sb) + bitof_(la) + 8;
}

long lz77_gpu_block_size(int sb)
{
    if (sb == -1) sb = LZ77_DEFAULT_SB;
    return sb <= 8191 ? 65536 : 524288;
}

int lz77_gpu_init(int device) { (void)device; return LZ77_OK; }
const char *lz77_gpu_last_error(void) { return ""; }
const char *lz77_gpu_strerror(int rc) { (void)rc; return "stand-in"; }
void *lz77_gpu_host_alloc(long n) { return malloc((size_t)n); }
void lz77_gpu_host_free(void *p) { free(p); }
long lz77_gpu_encode_bound(long n, int sb, int la) { return lz77o_encode_bound(n, sb, la); }

int lz77_gpu_encode(const unsigned char *in, long n, int sb, int la, unsigned char *out, long cap,
                    long *n_out)
{
    const int esb = sb == -1 ? LZ77_DEFAULT_SB : sb;
    const long r = lz77o_segmented_encode(in, n, sb, la, lz77_gpu_block_size(esb), 1024, out, cap, NULL);
    if (r == LZ77O_E_SPACE)
        return LZ77_E_SPACE;
    if (r < 0)
        return LZ77_E_ARG;
    *n_out = r;
    return LZ77_OK;
}

int lz77_gpu_decode_size(const unsigned char *in, long n, long *n_out)
{
    const long r = lz77o_decode(in, n, NULL, 0);
    if (r < 0)
        return LZ77_E_STREAM;
    *n_out = r;
    return LZ77_OK;
}

int lz77_gpu_decode(const unsigned char *in, long n, unsigned char *out, long cap, long *n_out)
{
    const long need = lz77o_decode(in, n, NULL, 0);
    if (need < 0)
        return LZ77_E_STREAM;
    *n_out = need;
    if (need > cap)
        return LZ77_E_SPACE;
    return lz77o_decode(in, n, out, cap) < 0 ? LZ77_E_STREAM : LZ77_OK;
}

int lz77_mgpu_init(int n_gpus) { (void)n_gpus; return LZ77_OK; }

int lz77_mgpu_encode(const unsigned char *in, long n, int sb, int la, unsigned char *out, long cap,
                     long *n_out)
{
    return lz77_gpu_encode(in, n, sb, la, out, cap, n_out);
}

int lz77_mgpu_decode(const unsigned char *in, long n, unsigned char *out, long cap, long *n_out)
{
    return lz77_gpu_decode(in, n, out, cap, n_out);
}
