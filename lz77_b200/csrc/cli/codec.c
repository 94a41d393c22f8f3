/*
 * codec.c -- encode()/decode() with the reference's signatures, running on the
 * GPU through liblz77b200.so.  The compressed side keeps the reference's
 * bitFILE handle shape; here it simply carries the FILE*, because whole
 * buffers (not single bits, bitio.c:203-298) cross the library boundary.
 *
 * Error behaviour follows the reference where it has one: a read error prints
 * a message and returns (lz77.c:79-82); failures that have no counterpart in
 * the reference (no GPU, CUDA error, malformed stream) print to stderr and
 * exit(EXIT_FAILURE) like lz77.c:273-277 does for a bit-read error.
 */
#include "codec.h"

#include <stdlib.h>
#include <string.h>

#include "lz77_b200.h"

struct bitFILE {
    FILE *file;
    int mode;
};

static int g_device = 0, g_gpus = 1;
static long g_piece_mib = 1024, g_out_mib = 4096;

void lz77_cli_set_device(int device) { g_device = device; }
void lz77_cli_set_gpus(int n) { g_gpus = n < 1 ? 1 : n; }
void lz77_cli_set_piece_mib(long mib) { g_piece_mib = mib < 1 ? 1 : mib; }
void lz77_cli_set_out_mib(long mib) { g_out_mib = mib < 1 ? 1 : mib; }

struct bitFILE *bitIO_open(const char *path, int mode)
{
    struct bitFILE *b;
    if (path == NULL || (mode != BIT_IO_W && mode != BIT_IO_R))
        return NULL;
    b = calloc(1, sizeof *b);
    if (b == NULL)
        return NULL;
    b->mode = mode;
    b->file = fopen(path, mode == BIT_IO_W ? "wb" : "rb");
    if (b->file == NULL) {
        free(b);
        return NULL;
    }
    return b;
}

int bitIO_close(struct bitFILE *b)
{
    int rc;
    if (b == NULL)
        return -1;
    rc = fclose(b->file);
    free(b);
    return rc == 0 ? 0 : -1;
}

static void die(const char *what, int rc)
{
    fprintf(stderr, "lz77: %s: %s", what, lz77_gpu_strerror(rc));
    if (rc == LZ77_E_CUDA)
        fprintf(stderr, " (%s)", lz77_gpu_last_error());
    fputc('\n', stderr);
    exit(EXIT_FAILURE);
}

static void bind_device(void)
{
    int rc = g_gpus > 1 ? lz77_mgpu_init(g_gpus) : lz77_gpu_init(g_device);
    if (rc != LZ77_OK)
        die("initialising the GPU", rc);
}

/* one library call: the bound device, or every device of -G */
static int codec_encode(const unsigned char *in, long n_in, int sb, int la, unsigned char *out,
                        long cap, long *n_out)
{
    if (g_gpus > 1)
        return lz77_mgpu_encode(in, n_in, sb, la, out, cap, n_out);
    return lz77_gpu_encode(in, n_in, sb, la, out, cap, n_out);
}

static int codec_decode(const unsigned char *in, long n_in, unsigned char *out, long cap,
                        long *n_out)
{
    if (g_gpus > 1)
        return lz77_mgpu_decode(in, n_in, out, cap, n_out);
    return lz77_gpu_decode(in, n_in, out, cap, n_out);
}

/* Bytes of input encoded per library call.  The encoder's blocks are independent,
 * so the stream of a large file is the streams of its pieces back to back: the
 * file never has to fit in (pinned) memory, like the reference's O(1)-memory
 * window loop (lz77.c:78,113-129).  A multiple of every block size. */
static long piece_bytes(void) { return g_piece_mib << 20; }

/* appends `nbits` bits (LSB-first, bitio.c:203-239) of src to the output file;
 * *acc / *acc_bits carry the bits of the last, still incomplete byte */
static int put_bits(FILE *f, const unsigned char *src, long nbits, unsigned *acc, int *acc_bits)
{
    long nbytes = nbits >> 3, i;
    int tail = (int)(nbits & 7);
    if (*acc_bits == 0) {
        if (nbytes > 0 && fwrite(src, 1, (size_t)nbytes, f) != (size_t)nbytes)
            return -1;
    } else {
        /* the stream position is inside a byte: shift everything by acc_bits */
        enum { BUF = 1 << 16 };
        static unsigned char buf[BUF];
        long done = 0;
        while (done < nbytes) {
            long n = nbytes - done < BUF ? nbytes - done : BUF;
            for (i = 0; i < n; i++) {
                unsigned v = *acc | ((unsigned)src[done + i] << *acc_bits);
                buf[i] = (unsigned char)v;
                *acc = v >> 8;
            }
            if (fwrite(buf, 1, (size_t)n, f) != (size_t)n)
                return -1;
            done += n;
        }
    }
    if (tail) {
        unsigned v = *acc | ((unsigned)(src[nbytes] & ((1u << tail) - 1u)) << *acc_bits);
        int total = *acc_bits + tail;
        if (total >= 8) {
            if (fputc((int)(v & 0xff), f) == EOF)
                return -1;
            v >>= 8;
            total -= 8;
        }
        *acc = v;
        *acc_bits = total;
    }
    return 0;
}

void encode(FILE *file, struct bitFILE *out, int la, int sb)
{
    const long piece = piece_bytes();
    const int esb = sb == -1 ? LZ77_DEFAULT_SB : sb, ela = la == -1 ? LZ77_DEFAULT_LA : la;
    const int tbits = lz77_token_bits(esb, ela);
    long cap = 0;
    unsigned char *in, *obuf = NULL;
    unsigned acc = 0;
    int acc_bits = 0, first = 1, rc;

    bind_device();
    in = lz77_gpu_host_alloc(piece);
    if (in == NULL)
        die("allocating the input buffer", LZ77_E_NOMEM);
    for (;;) {
        long n_out = 0, n_tokens;
        long n_in = (long)fread(in, 1, (size_t)piece, file);
        if (ferror(file)) {
            printf("Error loading the data in the window.\n"); /* lz77.c:79-82 */
            break;
        }
        if (n_in == 0 && !first)
            break;
        if (obuf == NULL) {
            cap = lz77_gpu_encode_bound(n_in < piece ? n_in : piece, sb, la) + 16;
            obuf = lz77_gpu_host_alloc(cap);
            if (obuf == NULL)
                die("allocating the output buffer", LZ77_E_NOMEM);
        }
        rc = codec_encode(in, n_in, sb, la, obuf, cap, &n_out);
        if (rc != LZ77_OK)
            die("encoding", rc);
        if (first && fwrite(obuf, 1, 4, out->file) != 4) /* header, lz77.c:74-75 */
            perror("Writing output file");
        /* padding is < 8 < T bits, so the token count follows from the size */
        n_tokens = ((n_out - 4) * 8) / tbits;
        if (put_bits(out->file, obuf + 4, n_tokens * tbits, &acc, &acc_bits) != 0)
            perror("Writing output file");
        first = 0;
        if (n_in < piece)
            break;
    }
    if (acc_bits > 0) /* zero padded last byte, bitio.c:180-182 */
        fputc((int)(acc & 0xff), out->file);
    lz77_gpu_host_free(in);
    if (obuf != NULL)
        lz77_gpu_host_free(obuf);
}

/* copies nbits bits from bit src_bit of src to bit dst_bit of dst (LSB-first bit
 * numbering, bitio.c:203-298); the bits of dst behind the copy must be zero */
static void copy_bits(unsigned char *dst, long dst_bit, const unsigned char *src, long src_bit,
                      long nbits)
{
    if (((dst_bit | src_bit) & 7) == 0) {
        memcpy(dst + (dst_bit >> 3), src + (src_bit >> 3), (size_t)(nbits >> 3));
        dst_bit += nbits & ~7L;
        src_bit += nbits & ~7L;
        nbits &= 7;
    }
    while (nbits > 0) {
        /* up to 8 bits that end at a source byte boundary */
        int s_off = (int)(src_bit & 7), d_off = (int)(dst_bit & 7);
        int take = 8 - s_off;
        unsigned v;
        if (take > nbits)
            take = (int)nbits;
        v = ((unsigned)src[src_bit >> 3] >> s_off) & ((1u << take) - 1u);
        dst[dst_bit >> 3] |= (unsigned char)(v << d_off);
        if (d_off + take > 8)
            dst[(dst_bit >> 3) + 1] |= (unsigned char)(v >> (8 - d_off));
        src_bit += take;
        dst_bit += take;
        nbits -= take;
    }
}

/*
 * The stream is decoded in pieces of tokens, so neither the compressed file nor
 * the output has to fit in memory (the reference's loop keeps SB bytes,
 * lz77.c:160-195).  Tokens are fixed width, so a piece is any run of whole tokens;
 * what a piece needs from its past is at most the last SB output bytes (lz77.c:184).
 * Every library call therefore gets a standalone stream: the header, the retained
 * output tail re-encoded as literal tokens (off 0, len 0, next = byte), then the
 * piece's tokens; the tail's bytes are dropped from the result.  The tail starts on a
 * block boundary of the output, so a stream of the block encoder stays aligned to
 * its blocks (and keeps decoding block-parallel).
 */
void decode(struct bitFILE *file, FILE *out)
{
    unsigned char hdr[4];
    unsigned char *raw, *sbuf = NULL, *obuf = NULL, *hist;
    long sbuf_cap = 0, obuf_cap = 0, hist_len = 0, raw_cap, piece_tokens;
    long out_limit = g_out_mib << 20;
    int sb, la, ob, lb, tbits, rc, last = 0;
    long block;

    bind_device();
    if (fread(hdr, 1, 4, file->file) != 4) {
        /* the reference reads garbage parameters from a short header and
         * produces an empty file; keep the empty output, flag nothing */
        return;
    }
    sb = hdr[0] | (hdr[1] << 8);
    la = hdr[2] | (hdr[3] << 8);
    if (sb < 1 || la < 1 || la > LZ77_MAX_LA)
        die("decoding", LZ77_E_STREAM);
    ob = lz77_bitof(sb);
    lb = lz77_bitof(la);
    tbits = ob + lb + 8;
    block = lz77_gpu_block_size(sb);
    /* whole tokens, a whole number of bytes */
    piece_tokens = ((piece_bytes() / 4) * 8 / tbits) & ~7L;
    if (piece_tokens < 8)
        piece_tokens = 8;
    raw_cap = piece_tokens / 8 * tbits;
    raw = malloc((size_t)raw_cap + 8);
    hist = malloc((size_t)(2 * block));
    if (raw == NULL || hist == NULL)
        die("allocating the input buffer", LZ77_E_NOMEM);

    while (!last) {
        long got = (long)fread(raw, 1, (size_t)raw_cap, file->file);
        long n_tok, cursor = 0;
        if (ferror(file->file)) {
            perror("Error reading bits"); /* lz77.c:273-277 */
            exit(EXIT_FAILURE);
        }
        last = got < raw_cap;
        /* lz77.c:271-280: a short read ends the stream, trailing bits < T are padding */
        n_tok = last ? (got * 8) / tbits : piece_tokens;
        while (cursor < n_tok) {
            long take = n_tok - cursor, n = 0, m = 0, need, bit, i;
            for (;;) {
                need = 4 + ((hist_len + take) * tbits + 7) / 8 + 16;
                if (need > sbuf_cap) {
                    if (sbuf != NULL)
                        lz77_gpu_host_free(sbuf);
                    sbuf_cap = need + need / 4;
                    sbuf = lz77_gpu_host_alloc(sbuf_cap);
                    if (sbuf == NULL)
                        die("allocating the input buffer", LZ77_E_NOMEM);
                }
                memset(sbuf, 0, (size_t)need);
                memcpy(sbuf, hdr, 4);
                bit = 32;
                for (i = 0; i < hist_len; i++) { /* the tail as literal tokens */
                    unsigned char v[5];
                    unsigned long long t = (unsigned long long)hist[i] << (ob + lb);
                    v[0] = (unsigned char)t, v[1] = (unsigned char)(t >> 8);
                    v[2] = (unsigned char)(t >> 16), v[3] = (unsigned char)(t >> 24);
                    v[4] = (unsigned char)(t >> 32);
                    copy_bits(sbuf, bit, v, 0, tbits);
                    bit += tbits;
                }
                copy_bits(sbuf, bit, raw, cursor * tbits, take * tbits);
                bit += take * tbits;
                rc = lz77_gpu_decode_size(sbuf, (bit + 7) / 8, &n);
                if (rc != LZ77_OK)
                    die("decoding", rc);
                if (n - hist_len <= out_limit || take <= 8)
                    break;
                take = (take / 2 + 7) & ~7L; /* highly compressible: smaller piece */
            }
            if (n + 16 > obuf_cap) {
                if (obuf != NULL)
                    lz77_gpu_host_free(obuf);
                obuf_cap = n + n / 4 + 16;
                obuf = lz77_gpu_host_alloc(obuf_cap);
                if (obuf == NULL)
                    die("allocating the output buffer", LZ77_E_NOMEM);
            }
            rc = codec_decode(sbuf, (bit + 7) / 8, obuf, obuf_cap, &m);
            if (rc != LZ77_OK)
                die("decoding", rc);
            if (m > hist_len &&
                fwrite(obuf + hist_len, 1, (size_t)(m - hist_len), out) != (size_t)(m - hist_len))
                perror("Writing output file");
            /* obuf starts on a block boundary of the output (or at its start): keep from
             * the last-but-one block boundary on, at least `block` > SB bytes */
            {
                long keep = (m % block) + block;
                if (keep > m)
                    keep = m;
                memcpy(hist, obuf + (m - keep), (size_t)keep);
                hist_len = keep;
            }
            cursor += take;
        }
    }
    free(raw);
    free(hist);
    if (sbuf != NULL)
        lz77_gpu_host_free(sbuf);
    if (obuf != NULL)
        lz77_gpu_host_free(obuf);
}
