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

static int g_device = 0;

void lz77_cli_set_device(int device) { g_device = device; }

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
    int rc = lz77_gpu_init(g_device);
    if (rc != LZ77_OK)
        die("initialising the GPU", rc);
}

/* read the rest of a stream into pinned memory; returns NULL on a read error */
static unsigned char *slurp(FILE *f, long *n_out)
{
    long cap = 1L << 24, n = 0;
    unsigned char *buf = lz77_gpu_host_alloc(cap);
    if (buf == NULL)
        die("allocating the input buffer", LZ77_E_NOMEM);
    for (;;) {
        size_t got = fread(buf + n, 1, (size_t)(cap - n), f);
        n += (long)got;
        if (n < cap)
            break;
        {
            unsigned char *bigger = lz77_gpu_host_alloc(cap * 2);
            if (bigger == NULL)
                die("allocating the input buffer", LZ77_E_NOMEM);
            memcpy(bigger, buf, (size_t)n);
            lz77_gpu_host_free(buf);
            buf = bigger;
            cap *= 2;
        }
    }
    if (ferror(f)) {
        lz77_gpu_host_free(buf);
        return NULL;
    }
    *n_out = n;
    return buf;
}

void encode(FILE *file, struct bitFILE *out, int la, int sb)
{
    long n_in = 0, n_out = 0, cap;
    unsigned char *in, *obuf;
    int rc;

    bind_device();
    in = slurp(file, &n_in);
    if (in == NULL) {
        printf("Error loading the data in the window.\n"); /* lz77.c:79-82 */
        return;
    }
    cap = lz77_gpu_encode_bound(n_in, sb, la) + 16;
    obuf = lz77_gpu_host_alloc(cap);
    if (obuf == NULL)
        die("allocating the output buffer", LZ77_E_NOMEM);
    rc = lz77_gpu_encode(in, n_in, sb, la, obuf, cap, &n_out);
    if (rc != LZ77_OK)
        die("encoding", rc);
    if (fwrite(obuf, 1, (size_t)n_out, out->file) != (size_t)n_out)
        perror("Writing output file");
    lz77_gpu_host_free(in);
    lz77_gpu_host_free(obuf);
}

void decode(struct bitFILE *file, FILE *out)
{
    long n_in = 0, n_out = 0, n = 0;
    unsigned char *in, *obuf;
    int rc;

    bind_device();
    in = slurp(file->file, &n_in);
    if (in == NULL) {
        perror("Error reading bits"); /* lz77.c:273-277 */
        exit(EXIT_FAILURE);
    }
    if (n_in < 4) {
        /* the reference reads garbage parameters from a short header and
         * produces an empty file; keep the empty output, flag nothing */
        lz77_gpu_host_free(in);
        return;
    }
    rc = lz77_gpu_decode_size(in, n_in, &n_out);
    if (rc != LZ77_OK)
        die("decoding", rc);
    obuf = lz77_gpu_host_alloc(n_out + 16);
    if (obuf == NULL)
        die("allocating the output buffer", LZ77_E_NOMEM);
    rc = lz77_gpu_decode(in, n_in, obuf, n_out, &n);
    if (rc != LZ77_OK)
        die("decoding", rc);
    if (fwrite(obuf, 1, (size_t)n, out) != (size_t)n)
        perror("Writing output file");
    lz77_gpu_host_free(in);
    lz77_gpu_host_free(obuf);
}
