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

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <time.h>
#include <unistd.h>

#include "lz77_b200.h"

struct bitFILE {
    FILE *file;
    int mode;
};

static int g_device = 0, g_gpus = 1, g_verbose = 0;
/* pieces: 32 MiB of input per encode call; 4 MiB of stream per decode call -- the decoder's
 * pinned buffers (stream piece + two output pieces) are a fixed cost of ~5 ms per 16 MiB
 * pinned, and a 256 MiB file is decoded before three 70 MiB buffers are even registered */
static long g_piece_mib = 32, g_dpiece_mib = 4, g_out_mib = 4096;

void lz77_cli_set_device(int device) { g_device = device; }
void lz77_cli_set_gpus(int n) { g_gpus = n < 1 ? 1 : n; }
void lz77_cli_set_piece_mib(long mib) { g_piece_mib = g_dpiece_mib = mib < 1 ? 1 : mib; }
void lz77_cli_set_out_mib(long mib) { g_out_mib = mib < 1 ? 1 : mib; }
void lz77_cli_set_verbose(int on) { g_verbose = on; }

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
    int rc;
    if (g_gpus == 1 && getenv("CUDA_VISIBLE_DEVICES") == NULL) {
        /* one GPU: do not make the CUDA runtime enumerate (and create state on) the other
         * devices of the box -- most of the start-up time of a short run */
        char dev[16];
        snprintf(dev, sizeof dev, "%d", g_device);
        setenv("CUDA_VISIBLE_DEVICES", dev, 1);
        g_device = 0;
    }
    rc = g_gpus > 1 ? lz77_mgpu_init(g_gpus) : lz77_gpu_init(g_device);
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

/*
 * The reference's loop interleaves file I/O and compute a few KiB at a time
 * (lz77.c:113-129 fread, bitio.c:80-93 fwrite).  Here whole pieces cross the library
 * boundary, so the overlap is explicit: a ring of pinned piece buffers, a reader thread
 * that fills them (fread), the calling thread that runs the GPU codec on the filled ones,
 * and a writer thread that drains the results (fwrite) -- piece k+1 is being read and
 * piece k-1 written while piece k is on the GPU.
 */
/* positional file I/O on `n_thr` threads: a tmpfs / page-cache copy runs at one core's
 * memcpy speed, several of them side by side at several times that */
struct pio {
    int fd, write;
    unsigned char *buf;
    long off, len, done;
    int failed;
};

static void *pio_main(void *arg)
{
    struct pio *p = arg;
    while (p->done < p->len) {
        ssize_t r = p->write ? pwrite(p->fd, p->buf + p->done, (size_t)(p->len - p->done), p->off + p->done)
                             : pread(p->fd, p->buf + p->done, (size_t)(p->len - p->done), p->off + p->done);
        if (r < 0) {
            p->failed = 1;
            break;
        }
        if (r == 0)
            break; /* end of file */
        p->done += r;
    }
    return NULL;
}

enum { PIO_THREADS = 4 };

/* reads / writes [off, off + len) of fd; returns the bytes transferred (short only at the
 * end of the file) or -1 */
static long pio_run(int fd, int write, unsigned char *buf, long off, long len)
{
    struct pio part[PIO_THREADS];
    pthread_t th[PIO_THREADS];
    long per = (len / PIO_THREADS + 4095) & ~4095L, total = 0;
    int i, n = 0, failed = 0;
    /* small transfers, and writes: one thread (writes to a fresh tmpfs / page-cache file are
     * bound by page allocation, which does not scale with threads: measured 2.3 GB/s on one
     * thread, 2.0 on four; reads go from 3.1 to 4-6 GB/s) */
    if (len < (4L << 20) || write)
        per = len;
    for (i = 0; i < PIO_THREADS && (long)i * per < len; i++, n++) {
        part[i].fd = fd, part[i].write = write, part[i].buf = buf + (long)i * per;
        part[i].off = off + (long)i * per;
        part[i].len = len - (long)i * per < per ? len - (long)i * per : per;
        part[i].done = 0, part[i].failed = 0;
        if (i > 0 && pthread_create(&th[i], NULL, pio_main, &part[i]) != 0) {
            pio_main(&part[i]);
            th[i] = 0;
        }
    }
    if (n > 0)
        pio_main(&part[0]);
    for (i = 1; i < n; i++)
        if (th[i])
            pthread_join(th[i], NULL);
    for (i = 0; i < n; i++) {
        failed |= part[i].failed;
        total += part[i].done;
        if (part[i].done < part[i].len)
            break; /* a short part: the file ended inside it */
    }
    return failed ? -1 : total;
}

enum { RING = 3 };

struct slot {
    unsigned char *in, *out; /* pinned */
    long n_in, n_out, cap;   /* bytes read / stream bytes produced (with its 4-byte header) / size of out */
    int eof;                 /* last piece of the input */
    int state;               /* 0 free, 1 filled, 2 encoded */
};

struct ring {
    struct slot s[RING];
    pthread_mutex_t mu;
    pthread_cond_t cv;
    FILE *fin, *fout;
    int fd_in, fd_out;       /* >= 0: positional I/O on several threads (regular files) */
    long in_off, out_off;
    long piece;
    int tbits, failed_read, failed_write, pieces_read_done;
    unsigned acc;
    int acc_bits, wrote_header;
};

static void ring_wait(struct ring *r, int i, int want)
{
    pthread_mutex_lock(&r->mu);
    while (r->s[i].state != want)
        pthread_cond_wait(&r->cv, &r->mu);
    pthread_mutex_unlock(&r->mu);
}

static void ring_set(struct ring *r, int i, int state)
{
    pthread_mutex_lock(&r->mu);
    r->s[i].state = state;
    pthread_cond_broadcast(&r->cv);
    pthread_mutex_unlock(&r->mu);
}

static void *reader_main(void *arg)
{
    struct ring *r = arg;
    int i = 0, first = 1;
    for (;;) {
        struct slot *sl = &r->s[i];
        ring_wait(r, i, 0);
        if (r->fd_in >= 0) {
            sl->n_in = pio_run(r->fd_in, 0, sl->in, r->in_off, r->piece);
            if (sl->n_in < 0) {
                r->failed_read = 1;
                sl->n_in = 0;
            }
            r->in_off += sl->n_in;
        } else {
            sl->n_in = (long)fread(sl->in, 1, (size_t)r->piece, r->fin);
            if (ferror(r->fin))
                r->failed_read = 1;
        }
        sl->eof = sl->n_in < r->piece || r->failed_read;
        if (sl->n_in == 0 && !first)
            sl->eof = 2; /* nothing left: no piece, just the end */
        first = 0;
        ring_set(r, i, 1);
        if (sl->eof)
            return NULL;
        i = (i + 1) % RING;
    }
}

static void *writer_main(void *arg)
{
    struct ring *r = arg;
    int i = 0;
    for (;;) {
        struct slot *sl = &r->s[i];
        int eof;
        ring_wait(r, i, 2);
        eof = sl->eof;
        if (sl->n_out > 0 && !r->failed_write) {
            /* padding is < 8 < T bits, so the token count follows from the size */
            long n_tokens = ((sl->n_out - 4) * 8) / r->tbits;
            if (r->fd_out >= 0) { /* byte-aligned tokens: payloads land at known offsets */
                long nb = n_tokens * r->tbits / 8;
                if (!r->wrote_header) {
                    if (pio_run(r->fd_out, 1, sl->out, 0, 4) != 4)
                        r->failed_write = 1;
                    r->out_off = 4;
                    r->wrote_header = 1;
                }
                if (pio_run(r->fd_out, 1, sl->out + 4, r->out_off, nb) != nb)
                    r->failed_write = 1;
                r->out_off += nb;
            } else {
                if (!r->wrote_header) { /* header, lz77.c:74-75 */
                    if (fwrite(sl->out, 1, 4, r->fout) != 4)
                        r->failed_write = 1;
                    r->wrote_header = 1;
                }
                if (put_bits(r->fout, sl->out + 4, n_tokens * r->tbits, &r->acc, &r->acc_bits) != 0)
                    r->failed_write = 1;
            }
        }
        ring_set(r, i, 0);
        if (eof)
            return NULL;
        i = (i + 1) % RING;
    }
}

void encode(FILE *file, struct bitFILE *out, int la, int sb)
{
    const int esb = sb == -1 ? LZ77_DEFAULT_SB : sb, ela = la == -1 ? LZ77_DEFAULT_LA : la;
    struct ring r;
    pthread_t reader, writer;
    struct timespec t0, t1;
    long cap, total_in = 0, total_out = 0;
    int i, rc;

    bind_device();
    memset(&r, 0, sizeof r);
    r.piece = piece_bytes();
    r.tbits = lz77_token_bits(esb, ela);
    r.fin = file;
    r.fout = out->file;
    r.fd_in = r.fd_out = -1;
    {
        /* regular files: positional reads / writes on several threads */
        struct stat st;
        long pos = ftell(file);
        if (pos >= 0 && fstat(fileno(file), &st) == 0 && S_ISREG(st.st_mode)) {
            r.fd_in = fileno(file);
            r.in_off = pos;
        }
        fflush(out->file);
        if ((r.tbits & 7) == 0 && ftell(out->file) == 0 && fstat(fileno(out->file), &st) == 0 &&
            S_ISREG(st.st_mode))
            r.fd_out = fileno(out->file);
    }
    pthread_mutex_init(&r.mu, NULL);
    pthread_cond_init(&r.cv, NULL);
    /* small files: no point in pinning three full pieces */
    {
        long pos = ftell(file), size = -1;
        if (pos >= 0 && fseek(file, 0, SEEK_END) == 0) {
            size = ftell(file) - pos;
            fseek(file, pos, SEEK_SET);
        }
        if (size >= 0 && size < r.piece) {
            const long block = 1L << 19; /* a multiple of every block size */
            r.piece = (size + block) / block * block;
        }
    }
    /* the worst case is T/8 bytes per input byte; text and logs need well under one, so
     * start there and fall back to the bound when a piece does not fit */
    cap = r.piece + r.piece / 4 + 65536;
    if (cap > lz77_gpu_encode_bound(r.piece, sb, la) + 16)
        cap = lz77_gpu_encode_bound(r.piece, sb, la) + 16;
    for (i = 0; i < RING; i++) {
        r.s[i].in = lz77_gpu_host_alloc(r.piece);
        r.s[i].out = lz77_gpu_host_alloc(cap);
        r.s[i].cap = cap;
        if (r.s[i].in == NULL || r.s[i].out == NULL)
            die("allocating the piece buffers", LZ77_E_NOMEM);
    }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_create(&reader, NULL, reader_main, &r);
    pthread_create(&writer, NULL, writer_main, &r);
    for (i = 0;; i = (i + 1) % RING) {
        struct slot *sl = &r.s[i];
        int eof;
        ring_wait(&r, i, 1);
        eof = sl->eof;
        sl->n_out = 0;
        if (r.failed_read) {
            printf("Error loading the data in the window.\n"); /* lz77.c:79-82 */
        } else if (eof != 2) {
            rc = codec_encode(sl->in, sl->n_in, sb, la, sl->out, sl->cap, &sl->n_out);
            if (rc == LZ77_E_SPACE) { /* incompressible: this slot gets the worst-case buffer */
                lz77_gpu_host_free(sl->out);
                sl->cap = lz77_gpu_encode_bound(r.piece, sb, la) + 16;
                sl->out = lz77_gpu_host_alloc(sl->cap);
                if (sl->out == NULL)
                    die("allocating the piece buffers", LZ77_E_NOMEM);
                rc = codec_encode(sl->in, sl->n_in, sb, la, sl->out, sl->cap, &sl->n_out);
            }
            if (rc != LZ77_OK)
                die("encoding", rc);
            total_in += sl->n_in;
            total_out += sl->n_out - 4;
        }
        ring_set(&r, i, 2);
        if (eof)
            break;
    }
    pthread_join(reader, NULL);
    pthread_join(writer, NULL);
    if (r.acc_bits > 0 && fputc((int)(r.acc & 0xff), out->file) == EOF) /* zero padded last byte, bitio.c:180-182 */
        r.failed_write = 1;
    if (fflush(out->file) != 0)
        r.failed_write = 1;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (g_verbose) {
        double s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        fprintf(stderr, "lz77: encoded %ld -> %ld bytes in %.3f s (%.2f GB/s, read + GPU + write "
                        "overlapped, %d GPU%s)\n",
                total_in, total_out + 4, s, s > 0 ? (double)total_in / s / 1e9 : 0.0, g_gpus,
                g_gpus > 1 ? "s" : "");
    }
    for (i = 0; i < RING; i++) {
        lz77_gpu_host_free(r.s[i].in);
        lz77_gpu_host_free(r.s[i].out);
    }
    pthread_mutex_destroy(&r.mu);
    pthread_cond_destroy(&r.cv);
    if (r.failed_write) { /* the reference swallows short writes (bitio.c:87-88); do not */
        perror("Writing output file");
        exit(EXIT_FAILURE);
    }
}

/* copies nbits bits from bit src_bit of src to bit dst_bit of dst (LSB-first bit
 * numbering, bitio.c:203-298); the bits of dst behind the copy must be zero.  src must
 * be readable for 8 bytes behind its last bit. */
static void copy_bits(unsigned char *dst, long dst_bit, const unsigned char *src, long src_bit,
                      long nbits)
{
    if (((dst_bit ^ src_bit) & 7) == 0) {
        /* same position inside the byte: up to 7 head bits, then whole bytes */
        while (nbits > 0 && (dst_bit & 7) != 0) {
            dst[dst_bit >> 3] |= (unsigned char)(((src[src_bit >> 3] >> (src_bit & 7)) & 1u) << (dst_bit & 7));
            dst_bit++, src_bit++, nbits--;
        }
        memcpy(dst + (dst_bit >> 3), src + (src_bit >> 3), (size_t)(nbits >> 3));
        dst_bit += nbits & ~7L;
        src_bit += nbits & ~7L;
        nbits &= 7;
    } else {
        /* bring dst to a byte boundary, then 8 destination bytes per step */
        while (nbits > 0 && (dst_bit & 7) != 0) {
            dst[dst_bit >> 3] |= (unsigned char)(((src[src_bit >> 3] >> (src_bit & 7)) & 1u) << (dst_bit & 7));
            dst_bit++, src_bit++, nbits--;
        }
        while (nbits >= 64) {
            const int sh = (int)(src_bit & 7); /* != 0 here */
            unsigned long long lo, hi;
            memcpy(&lo, src + (src_bit >> 3), 8);
            hi = src[(src_bit >> 3) + 8];
            lo = (lo >> sh) | (hi << (64 - sh));
            memcpy(dst + (dst_bit >> 3), &lo, 8);
            dst_bit += 64, src_bit += 64, nbits -= 64;
        }
    }
    while (nbits > 0) { /* the last < 64 bits */
        dst[dst_bit >> 3] |= (unsigned char)(((src[src_bit >> 3] >> (src_bit & 7)) & 1u) << (dst_bit & 7));
        dst_bit++, src_bit++, nbits--;
    }
}

/* n literal tokens (off 0, len 0, next = byte; lz77.c:249-251) from bit `bit` of dst on */
static void put_literals(unsigned char *dst, long bit, const unsigned char *bytes, long n,
                         int tbits, int lit_shift)
{
    long i;
    if ((tbits & 7) == 0 && (bit & 7) == 0) {
        unsigned char *p = dst + (bit >> 3) + (lit_shift >> 3); /* lit_shift = tbits - 8 */
        const int step = tbits >> 3;
        for (i = 0; i < n; i++, p += step)
            *p = bytes[i];
        return;
    }
    for (i = 0; i < n; i++) {
        unsigned long long t = (unsigned long long)bytes[i] << lit_shift;
        unsigned char v[8];
        memcpy(v, &t, 8);
        copy_bits(dst, bit, v, 0, tbits);
        bit += tbits;
    }
}

/* the writer side of decode(): one pending write at a time, on its own thread */
struct wjob {
    pthread_mutex_t mu;
    pthread_cond_t cv;
    FILE *f;
    int fd;      /* >= 0: positional writes on several threads (a regular file) */
    long off;
    const unsigned char *p;
    long n;
    int pending, quit, failed;
};

static void *wjob_main(void *arg)
{
    struct wjob *w = arg;
    pthread_mutex_lock(&w->mu);
    for (;;) {
        while (!w->pending && !w->quit)
            pthread_cond_wait(&w->cv, &w->mu);
        if (!w->pending && w->quit)
            break;
        pthread_mutex_unlock(&w->mu);
        if (w->n > 0 && w->fd >= 0) {
            if (pio_run(w->fd, 1, (unsigned char *)w->p, w->off, w->n) != w->n)
                w->failed = 1;
            w->off += w->n;
        } else if (w->n > 0 && fwrite(w->p, 1, (size_t)w->n, w->f) != (size_t)w->n) {
            w->failed = 1;
        }
        pthread_mutex_lock(&w->mu);
        w->pending = 0;
        pthread_cond_broadcast(&w->cv);
    }
    pthread_mutex_unlock(&w->mu);
    return NULL;
}

static void wjob_wait(struct wjob *w)
{
    pthread_mutex_lock(&w->mu);
    while (w->pending)
        pthread_cond_wait(&w->cv, &w->mu);
    pthread_mutex_unlock(&w->mu);
}

static void wjob_submit(struct wjob *w, const unsigned char *p, long n)
{
    pthread_mutex_lock(&w->mu);
    while (w->pending)
        pthread_cond_wait(&w->cv, &w->mu);
    w->p = p;
    w->n = n;
    w->pending = 1;
    pthread_cond_broadcast(&w->cv);
    pthread_mutex_unlock(&w->mu);
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
 * its blocks (and keeps decoding block-parallel).  The decoded piece is written by a
 * second thread while the next piece is on the GPU (two output buffers).
 */
/* the reader side of decode(): the next piece of the stream, read ahead on its own thread */
struct dread {
    pthread_mutex_t mu;
    pthread_cond_t cv;
    FILE *f;
    unsigned char *buf[2];
    long off, cap; /* a piece lands at buf[i] + off, cap bytes at most */
    long got[2];
    int full[2], err;
};

static void *dread_main(void *arg)
{
    struct dread *d = arg;
    int i = 0;
    for (;;) {
        long got;
        int err;
        pthread_mutex_lock(&d->mu);
        while (d->full[i])
            pthread_cond_wait(&d->cv, &d->mu);
        pthread_mutex_unlock(&d->mu);
        got = (long)fread(d->buf[i] + d->off, 1, (size_t)d->cap, d->f);
        err = ferror(d->f);
        memset(d->buf[i] + d->off + got, 0, 16);
        pthread_mutex_lock(&d->mu);
        d->got[i] = got;
        d->err |= err;
        d->full[i] = 1;
        pthread_cond_broadcast(&d->cv);
        pthread_mutex_unlock(&d->mu);
        if (got < d->cap || err)
            break; /* lz77.c:271-280: a short read ends the stream */
        i ^= 1;
    }
    return NULL;
}

void decode(struct bitFILE *file, FILE *out)
{
    unsigned char hdr[4];
    unsigned char *sbuf = NULL, *obuf[2] = {NULL, NULL}, *hist;
    long sbuf_cap = 0, obuf_cap[2] = {0, 0}, hist_len = 0, raw_cap, piece_tokens, lead;
    long out_limit = g_out_mib << 20, total_in = 4, total_out = 0;
    int sb, la, ob, lb, tbits, rc, last = 0, cur = 0, slot = 0, aligned;
    long block;
    struct wjob w;
    struct dread rd;
    pthread_t writer, reader;
    struct timespec t0, t1;

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
    aligned = (tbits & 7) == 0;
    block = lz77_gpu_block_size(sb);
    /* whole tokens, a whole number of bytes */
    piece_tokens = ((g_dpiece_mib << 20) * 8 / tbits) & ~7L;
    if (piece_tokens < 8)
        piece_tokens = 8;
    raw_cap = piece_tokens / 8 * tbits;
    hist = malloc((size_t)(2 * block));
    if (hist == NULL)
        die("allocating the input buffer", LZ77_E_NOMEM);
    /* Byte-aligned tokens (both benchmark parameter sets): a piece is read straight into
     * pinned memory behind a lead area, and the standalone stream of a library call --
     * header, the retained tail as literal tokens, the piece's tokens -- is completed in
     * place in FRONT of the tokens (the lead area, or tokens this piece has consumed):
     * nothing is copied.  Other widths are assembled bit by bit in a second buffer. */
    lead = aligned ? ((4 + 2 * block * (tbits >> 3) + 63) & ~63L) : 0;
    memset(&rd, 0, sizeof rd);
    pthread_mutex_init(&rd.mu, NULL);
    pthread_cond_init(&rd.cv, NULL);
    rd.f = file->file;
    rd.off = lead;
    rd.cap = raw_cap;
    for (rc = 0; rc < 2; rc++) {
        rd.buf[rc] = aligned ? lz77_gpu_host_alloc(lead + raw_cap + 64) : malloc((size_t)raw_cap + 64);
        if (rd.buf[rc] == NULL)
            die("allocating the input buffer", LZ77_E_NOMEM);
    }
    memset(&w, 0, sizeof w);
    pthread_mutex_init(&w.mu, NULL);
    pthread_cond_init(&w.cv, NULL);
    w.f = out;
    w.fd = -1;
    {
        struct stat st;
        fflush(out);
        if (ftell(out) == 0 && fstat(fileno(out), &st) == 0 && S_ISREG(st.st_mode))
            w.fd = fileno(out);
    }
    pthread_create(&writer, NULL, wjob_main, &w);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_create(&reader, NULL, dread_main, &rd);

    while (!last) {
        unsigned char *raw;
        long got, n_tok, cursor = 0;
        pthread_mutex_lock(&rd.mu);
        while (!rd.full[slot])
            pthread_cond_wait(&rd.cv, &rd.mu);
        got = rd.got[slot];
        rc = rd.err;
        pthread_mutex_unlock(&rd.mu);
        if (rc) {
            perror("Error reading bits"); /* lz77.c:273-277 */
            exit(EXIT_FAILURE);
        }
        raw = rd.buf[slot] + lead;
        total_in += got;
        last = got < raw_cap;
        /* lz77.c:271-280: a short read ends the stream, trailing bits < T are padding */
        n_tok = last ? (got * 8) / tbits : piece_tokens;
        while (cursor < n_tok) {
            long take = n_tok - cursor, m = 0, n_stream;
            const unsigned char *stream;
            for (;;) {
                if (aligned) {
                    const long tb = tbits >> 3;
                    unsigned char *sp = raw + cursor * tb - hist_len * tb - 4;
                    memcpy(sp, hdr, 4);
                    memset(sp + 4, 0, (size_t)(hist_len * tb)); /* literal tokens: off 0, len 0 */
                    put_literals(sp, 32, hist, hist_len, tbits, ob + lb);
                    stream = sp;
                    n_stream = 4 + (hist_len + take) * tb;
                } else {
                    const long need = 4 + ((hist_len + take) * tbits + 7) / 8 + 32;
                    long bit = 32 + hist_len * tbits;
                    if (need > sbuf_cap) {
                        if (sbuf != NULL)
                            lz77_gpu_host_free(sbuf);
                        sbuf_cap = need + need / 4;
                        sbuf = lz77_gpu_host_alloc(sbuf_cap);
                        if (sbuf == NULL)
                            die("allocating the input buffer", LZ77_E_NOMEM);
                    }
                    /* header, the tail as literal tokens, the piece's tokens */
                    memcpy(sbuf, hdr, 4);
                    memset(sbuf + 4, 0, (size_t)(need - 4)); /* bits are OR-ed in */
                    put_literals(sbuf, 32, hist, hist_len, tbits, ob + lb);
                    copy_bits(sbuf, bit, raw, cursor * tbits, take * tbits);
                    bit += take * tbits;
                    stream = sbuf;
                    n_stream = (bit + 7) / 8;
                }
                if (obuf[cur] == NULL) {
                    /* first use of this buffer: size it from the piece's decoded size (one
                     * token scan) -- pinning memory twice costs far more */
                    long want = 0;
                    rc = lz77_gpu_decode_size(stream, n_stream, &want);
                    if (rc != LZ77_OK)
                        die("decoding", rc);
                    obuf_cap[cur] = want + want / 4 + (1L << 20);
                    if (obuf_cap[cur] > out_limit + hist_len + (1L << 20))
                        obuf_cap[cur] = out_limit + hist_len + (1L << 20);
                    obuf[cur] = lz77_gpu_host_alloc(obuf_cap[cur]);
                    if (obuf[cur] == NULL)
                        die("allocating the output buffer", LZ77_E_NOMEM);
                }
                rc = codec_decode(stream, n_stream, obuf[cur], obuf_cap[cur] - 16, &m);
                if (rc == LZ77_OK)
                    break;
                if (rc != LZ77_E_SPACE)
                    die("decoding", rc);
                /* m = the decoded size: a bigger buffer, or -- highly compressible input --
                 * a smaller piece */
                if (m - hist_len <= out_limit || take <= 8) {
                    lz77_gpu_host_free(obuf[cur]);
                    obuf_cap[cur] = m + m / 8 + 16;
                    obuf[cur] = lz77_gpu_host_alloc(obuf_cap[cur]);
                    if (obuf[cur] == NULL)
                        die("allocating the output buffer", LZ77_E_NOMEM);
                } else {
                    take = (take / 2 + 7) & ~7L;
                }
            }
            /* obuf starts on a block boundary of the output (or at its start): keep from
             * the last-but-one block boundary on, at least `block` > SB bytes */
            {
                const long old_hist = hist_len;
                long keep = (m % block) + block;
                if (keep > m)
                    keep = m;
                memcpy(hist, obuf[cur] + (m - keep), (size_t)keep);
                hist_len = keep;
                wjob_submit(&w, obuf[cur] + old_hist, m - old_hist);
                total_out += m - old_hist;
            }
            cur ^= 1;
            cursor += take;
        }
        pthread_mutex_lock(&rd.mu); /* the piece is consumed: its buffer may be read into again */
        rd.full[slot] = 0;
        pthread_cond_broadcast(&rd.cv);
        pthread_mutex_unlock(&rd.mu);
        slot ^= 1;
    }
    pthread_join(reader, NULL);
    wjob_wait(&w);
    pthread_mutex_lock(&w.mu);
    w.quit = 1;
    pthread_cond_broadcast(&w.cv);
    pthread_mutex_unlock(&w.mu);
    pthread_join(writer, NULL);
    if (fflush(out) != 0)
        w.failed = 1;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (g_verbose) {
        double s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        fprintf(stderr, "lz77: decoded %ld -> %ld bytes in %.3f s (%.2f GB/s, read + GPU + write "
                        "overlapped, %d GPU%s)\n",
                total_in, total_out, s, s > 0 ? (double)total_out / s / 1e9 : 0.0, g_gpus,
                g_gpus > 1 ? "s" : "");
    }
    free(hist);
    for (rc = 0; rc < 2; rc++) {
        if (aligned)
            lz77_gpu_host_free(rd.buf[rc]);
        else
            free(rd.buf[rc]);
    }
    if (sbuf != NULL)
        lz77_gpu_host_free(sbuf);
    if (obuf[0] != NULL)
        lz77_gpu_host_free(obuf[0]);
    if (obuf[1] != NULL)
        lz77_gpu_host_free(obuf[1]);
    pthread_mutex_destroy(&w.mu);
    pthread_cond_destroy(&w.cv);
    pthread_mutex_destroy(&rd.mu);
    pthread_cond_destroy(&rd.cv);
    if (w.failed) {
        perror("Writing output file");
        exit(EXIT_FAILURE);
    }
}
