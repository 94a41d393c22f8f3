/*
 * lz77_oracle.c -- CPU restatement of the cstdvd/lz77 codec hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see lz77_oracle.h).  Parity status: PINNED against
 * the compiled reference and the Appendix-C known-answer vectors.
 *
 * Citations "ref:" are file:line into /root/reference.
 */
#include "lz77_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* bit widths                                                               */
/* ------------------------------------------------------------------------ */

/* ref: bitio.c:41-43  bitof(n) = (int)ceil(log(n)/log(2)).  For 1 <= n <= 65535
 * this equals the exact integer ceil(log2 n) (probed). */
int lz77o_bitof(int n)
{
    int b = 0;
    if (n <= 1)
        return 0;
    while ((1L << b) < (long)n)
        b++;
    return b;
}

/* ref: lz77.c:249-251 -- offset field, length field, 8-bit literal */
int lz77o_token_bits(int sb, int la)
{
    return lz77o_bitof(sb) + lz77o_bitof(la) + 8;
}

long lz77o_encode_bound(long n_in, int sb, int la)
{
    long t = lz77o_token_bits(sb, la);
    return 4 + (n_in * t + 7) / 8;
}

/* ------------------------------------------------------------------------ */
/* LSB-first bit sink / source over a memory buffer                         */
/* ------------------------------------------------------------------------ */

/* ref: bitio.c:213-229 -- value bit i goes to stream bit (pos+i); stream bit b
 * is bit (b & 7) of byte (b >> 3).  The buffer is pre-zeroed (bitio.c:92,143)
 * so the final partial byte is zero padded (bitio.c:180-182). */
typedef struct {
    uint8_t *buf;
    long cap;      /* bytes */
    long bitpos;   /* next bit to write */
    int overflow;
} bitsink;

static void sink_put(bitsink *s, uint32_t v, int nbits)
{
    for (int i = 0; i < nbits; i++) {
        long byte = s->bitpos >> 3;
        if (byte >= s->cap) {
            s->overflow = 1;
            return;
        }
        if ((v >> i) & 1u)
            s->buf[byte] |= (uint8_t)(1u << (s->bitpos & 7));
        s->bitpos++;
    }
}

typedef struct {
    const uint8_t *buf;
    long nbits;    /* total bits available */
    long bitpos;
} bitsource;

/* ref: bitio.c:256-298 -- reads up to nbits, stops at end of data and returns
 * how many bits it actually delivered (short count == EOF, lz77.c:271-280). */
static int source_get(bitsource *s, uint32_t *v, int nbits)
{
    uint32_t acc = 0;
    int got = 0;
    while (got < nbits && s->bitpos < s->nbits) {
        uint32_t bit = (s->buf[s->bitpos >> 3] >> (s->bitpos & 7)) & 1u;
        acc |= bit << got;
        s->bitpos++;
        got++;
    }
    *v = acc;
    return got;
}

/* ------------------------------------------------------------------------ */
/* memory "FILE" with fread/feof semantics                                  */
/* ------------------------------------------------------------------------ */

typedef struct {
    const uint8_t *data;
    long size, pos;
    int eof; /* set once a read came back short, like stdio's indicator */
} memfile;

static long mf_read(memfile *f, uint8_t *dst, long want)
{
    long left = f->size - f->pos;
    long got = want < left ? want : left;
    if (got > 0) {
        memcpy(dst, f->data + f->pos, (size_t)got);
        f->pos += got;
    }
    if (got < want)
        f->eof = 1;
    return got;
}

/* ------------------------------------------------------------------------ */
/* the reference's array-backed binary search tree                          */
/* ------------------------------------------------------------------------ */

/* ref: tree.c:23-27 -- node {len, off, parent, left, right}; one slot per
 * search-buffer position, slot = absolute offset mod SB (tree.c:66,187).
 * calloc'ed to zero (tree.c:37). */
typedef struct {
    int *off, *len, *parent, *left, *right;
    int root;
    int slots;
} bst;

static int bst_init(bst *t, int slots)
{
    t->slots = slots;
    t->root = -1;
    t->off = calloc((size_t)slots, sizeof(int));
    t->len = calloc((size_t)slots, sizeof(int));
    t->parent = calloc((size_t)slots, sizeof(int));
    t->left = calloc((size_t)slots, sizeof(int));
    t->right = calloc((size_t)slots, sizeof(int));
    return t->off && t->len && t->parent && t->left && t->right;
}

static void bst_free(bst *t)
{
    free(t->off);
    free(t->len);
    free(t->parent);
    free(t->left);
    free(t->right);
}

/* ref: tree.c:62-106.  Key = the `len` bytes at window[abs_off]; strictly
 * smaller goes left, everything else (including equal) goes right. */
static void bst_insert(bst *t, const uint8_t *win, int abs_off, int len)
{
    int slot = abs_off % t->slots;

    if (t->root == -1) {
        t->root = slot;
        t->parent[slot] = -1;
    } else {
        int at = t->root;
        for (;;) {
            int below;
            int goes_left =
                memcmp(win + abs_off, win + t->off[at], (size_t)len) < 0;
            below = goes_left ? t->left[at] : t->right[at];
            if (below == -1) {
                if (goes_left)
                    t->left[at] = slot;
                else
                    t->right[at] = slot;
                t->parent[slot] = at;
                break;
            }
            at = below;
        }
    }
    t->off[slot] = abs_off;
    t->len[slot] = len;
    t->left[slot] = -1;
    t->right[slot] = -1;
}

/* ref: tree.c:118-152.  Walk one root-to-leaf path; at each node count the
 * common prefix (capped at size-1), keep the first strictly longer one, then
 * branch on the first differing byte; stop on equality or a missing child. */
static void bst_find(const bst *t, const uint8_t *win, int index, int size,
                     int *best_off, int *best_len)
{
    *best_off = 0;
    *best_len = 0;
    if (t->root == -1)
        return;

    int at = t->root;
    for (;;) {
        const uint8_t *a = win + index;
        const uint8_t *b = win + t->off[at];
        int i = 0;
        while (a[i] == b[i] && i < size - 1)
            i++;
        if (i > *best_len) {
            *best_off = index - t->off[at];
            *best_len = i;
        }
        if (a[i] < b[i] && t->left[at] != -1)
            at = t->left[at];
        else if (a[i] > b[i] && t->right[at] != -1)
            at = t->right[at];
        else
            break;
    }
}

/* ref: tree.c:162-170 */
static int bst_leftmost(const bst *t, int at)
{
    while (t->left[at] != -1)
        at = t->left[at];
    return at;
}

/* ref: tree.c:182-243.  Standard BST removal of the slot holding absolute
 * offset abs_sb; a two-child node is replaced by the leftmost node of its
 * right subtree. */
static void bst_delete(bst *t, int abs_sb)
{
    int sb = abs_sb % t->slots;
    int up, repl;

    if (t->left[sb] == -1) {
        repl = t->right[sb];
        if (repl != -1)
            t->parent[repl] = t->parent[sb];
        up = t->parent[sb];
    } else if (t->right[sb] == -1) {
        repl = t->left[sb];
        t->parent[repl] = t->parent[sb];
        up = t->parent[sb];
    } else {
        repl = bst_leftmost(t, t->right[sb]);
        if (t->parent[repl] == sb) {
            up = t->parent[sb];
            t->parent[repl] = up;
        } else {
            int rp = t->parent[repl];
            t->left[rp] = t->right[repl];
            if (t->right[repl] != -1)
                t->parent[t->right[repl]] = rp;

            t->right[repl] = t->right[sb];
            t->parent[repl] = t->parent[sb];
            if (t->right[repl] != -1)
                t->parent[t->right[repl]] = repl;
            up = t->parent[repl];
        }
        t->left[repl] = t->left[sb];
        if (t->left[repl] != -1)
            t->parent[t->left[repl]] = repl;
    }

    if (up != -1) {
        if (t->right[up] == sb)
            t->right[up] = repl;
        else
            t->left[up] = repl;
    } else {
        t->root = repl;
    }
}

/* ref: tree.c:252-260.  Slots start at 0 (calloc) so the != -1 guard never
 * skips anything; every slot is shifted. */
static void bst_shift(bst *t, int by)
{
    for (int i = 0; i < t->slots; i++)
        if (t->off[i] != -1)
            t->off[i] -= by;
}

/* ------------------------------------------------------------------------ */
/* reference encoder                                                        */
/* ------------------------------------------------------------------------ */

/* ref: lz77.c:51-140 */
long lz77o_ref_encode(const uint8_t *in, long n_in, int sb, int la,
                      uint8_t *out, long out_cap)
{
    const int SB = (sb == -1) ? LZ77O_DEFAULT_SB : sb; /* lz77.c:65-66 */
    const int LA = (la == -1) ? LZ77O_DEFAULT_LA : la;
    if (SB < 1 || SB > 65535 || LA < 1 || LA > 65535 || n_in < 0)
        return LZ77O_E_ARG; /* SB == 0 divides by zero in the reference */

    const long WIN = 3L * SB + LA; /* lz77.c:67, N == 3 */
    const int off_bits = lz77o_bitof(SB);
    const int len_bits = lz77o_bitof(LA);

    uint8_t *win = calloc((size_t)WIN + 1, 1);
    bst tree;
    if (!win || !bst_init(&tree, SB)) {
        free(win);
        return LZ77O_E_NOMEM;
    }

    memset(out, 0, (size_t)out_cap);
    bitsink sink = { out, out_cap, 0, 0 };
    sink_put(&sink, (uint32_t)SB, 16); /* lz77.c:74-75, MAX_BIT_BUFFER 16 */
    sink_put(&sink, (uint32_t)LA, 16);

    memfile f = { in, n_in, 0, 0 };
    long pending = mf_read(&f, win, WIN); /* lz77.c:78: bytes not yet coded */
    int eof = f.eof;                       /* lz77.c:84 */

    int sb_len = 0;   /* bytes currently in the search buffer      */
    int sb_at = 0;    /* window index of the oldest search byte    */
    int la_at = 0;    /* window index of the first lookahead byte  */
    int la_len = pending > LA ? LA : (int)pending; /* lz77.c:87 */

    while (pending > 0) { /* lz77.c:89 */
        int m_off, m_len;
        bst_find(&tree, win, la_at, la_len, &m_off, &m_len); /* lz77.c:92,216 */
        uint8_t lit = win[la_at + m_len];                    /* lz77.c:221 */

        sink_put(&sink, (uint32_t)m_off, off_bits); /* lz77.c:249-251 */
        sink_put(&sink, (uint32_t)m_len, len_bits);
        sink_put(&sink, lit, 8);

        for (int i = 0; i < m_len + 1; i++) { /* lz77.c:98 */
            if (sb_len == SB) {               /* lz77.c:101-105 */
                bst_delete(&tree, sb_at);
                sb_at++;
            } else {
                sb_len++;
            }
            bst_insert(&tree, win, la_at, la_len); /* lz77.c:108 */
            la_at++;

            if (!eof && sb_at == 2 * SB) { /* lz77.c:111-129 */
                memmove(win, win + sb_at, (size_t)(sb_len + la_len));
                bst_shift(&tree, sb_at);
                sb_at = 0;
                la_at = sb_len;
                pending += mf_read(&f, win + sb_len + la_len,
                                   WIN - (sb_len + la_len));
                eof = f.eof;
            }
            pending--;                                   /* lz77.c:132 */
            la_len = pending > LA ? LA : (int)pending;   /* lz77.c:134 */
        }
    }

    bst_free(&tree);
    free(win);
    if (sink.overflow)
        return LZ77O_E_SPACE;
    return (sink.bitpos + 7) >> 3; /* bitio.c:180-182 flushes the partial byte */
}

/* ------------------------------------------------------------------------ */
/* decoder                                                                  */
/* ------------------------------------------------------------------------ */

static int read_header(const uint8_t *in, long n_in, int *sb, int *la)
{
    if (n_in < 4)
        return LZ77O_E_STREAM;
    *sb = in[0] | (in[1] << 8); /* lz77.c:157-158: two 16-bit LSB-first fields */
    *la = in[2] | (in[3] << 8);
    if (*sb < 1 || *la < 1)
        return LZ77O_E_STREAM; /* bitof(0) is undefined in the reference */
    return 0;
}

/* ref: lz77.c:148-197.  The reference keeps a 3*SB+LA byte buffer and slides
 * the last SB bytes to its front when it would overflow (lz77.c:172-175); the
 * same buffer mechanics are kept here so a back reference reads exactly the
 * byte the reference would read. */
long lz77o_decode(const uint8_t *in, long n_in, uint8_t *out, long out_cap)
{
    int SB, LA;
    int rc = read_header(in, n_in, &SB, &LA);
    if (rc)
        return rc;

    const int off_bits = lz77o_bitof(SB);
    const int len_bits = lz77o_bitof(LA);
    const int tok_bits = off_bits + len_bits + 8;
    const long WIN = 3L * SB + LA;

    uint8_t *buf = calloc((size_t)WIN + 65536 + 1, 1);
    if (!buf)
        return LZ77O_E_NOMEM;

    bitsource src = { in, n_in * 8, 32 };
    long produced = 0;
    long back = 0;

    for (;;) {
        uint32_t off, len, lit;
        int got = source_get(&src, &off, off_bits); /* lz77.c:266-268 */
        got += source_get(&src, &len, len_bits);
        got += source_get(&src, &lit, 8);
        if (got < tok_bits) /* lz77.c:271-280: short read == end of stream */
            break;

        if (back + (long)len > WIN - 1) { /* lz77.c:172-175 */
            memcpy(buf, buf + back - SB, (size_t)SB);
            back = SB;
        }
        if (len > 0 && (long)off > back) {
            /* the reference would index before its buffer here (UB) */
            free(buf);
            return LZ77O_E_STREAM;
        }
        for (uint32_t k = 0; k < len; k++) { /* lz77.c:178-188 */
            uint8_t c = buf[back - off];
            buf[back] = c;
            if (out) {
                if (produced >= out_cap) {
                    free(buf);
                    return LZ77O_E_SPACE;
                }
                out[produced] = c;
            }
            produced++;
            back++;
        }
        buf[back] = (uint8_t)lit; /* lz77.c:189-194 */
        if (out) {
            if (produced >= out_cap) {
                free(buf);
                return LZ77O_E_SPACE;
            }
            out[produced] = (uint8_t)lit;
        }
        produced++;
        back++;
    }
    free(buf);
    return produced;
}

long lz77o_unpack_tokens(const uint8_t *in, long n_in, int *sb, int *la,
                         int32_t *off, int32_t *len, uint8_t *next, long cap)
{
    int SB, LA;
    int rc = read_header(in, n_in, &SB, &LA);
    if (rc)
        return rc;
    if (sb)
        *sb = SB;
    if (la)
        *la = LA;
    const int off_bits = lz77o_bitof(SB);
    const int len_bits = lz77o_bitof(LA);
    const int tok_bits = off_bits + len_bits + 8;
    bitsource src = { in, n_in * 8, 32 };
    long k = 0;
    for (;;) {
        uint32_t o, l, c;
        int got = source_get(&src, &o, off_bits);
        got += source_get(&src, &l, len_bits);
        got += source_get(&src, &c, 8);
        if (got < tok_bits)
            break;
        if (k < cap) {
            if (off)
                off[k] = (int32_t)o;
            if (len)
                len[k] = (int32_t)l;
            if (next)
                next[k] = (uint8_t)c;
        }
        k++;
    }
    return k;
}

/* ------------------------------------------------------------------------ */
/* block-parallel encoder specification                                     */
/* ------------------------------------------------------------------------ */

/* Longest match for the lookahead at blk[p] against starts p-reach .. p-1,
 * OLDEST first, strictly longer wins => the farthest offset among the longest
 * (the reference's BST also tends to answer with the oldest suffix: its root
 * is the oldest node, see the 1,16,31.. offsets of SURVEY.md Appendix C).  The
 * choice among equal-length candidates does not change the token count; the
 * oldest one keeps the decoder's dependency chains short.  Capped at max_len.
 * Candidate strings may run into the lookahead, as in tree.c:136 where
 * window[node.off + i] has no upper clamp. */
static void longest_match(const uint8_t *blk, long p, long reach, int max_len,
                          int *m_off, int *m_len)
{
    int best = 0, best_off = 0;
    if (max_len > 0) {
        const uint8_t *a = blk + p;
        for (long d = reach; d >= 1; d--) {
            const uint8_t *b = a - d;
            if (b[best] != a[best] || b[0] != a[0])
                continue;
            int i = 0;
            while (i < max_len && a[i] == b[i])
                i++;
            if (i > best) {
                best = i;
                best_off = (int)d;
                if (best == max_len)
                    break;
            }
        }
    }
    *m_off = best_off;
    *m_len = best;
}

long lz77o_segmented_encode(const uint8_t *in, long n_in, int sb, int la,
                            long block, long segment, uint8_t *out,
                            long out_cap, long *n_tokens)
{
    const int SB = (sb == -1) ? LZ77O_DEFAULT_SB : sb;
    const int LA = (la == -1) ? LZ77O_DEFAULT_LA : la;
    if (SB < 1 || SB > 65535 || LA < 1 || LA > 65535 || n_in < 0)
        return LZ77O_E_ARG;
    /* block <= 0: ONE block, i.e. the reference's sliding window over the whole input
     * (lz77.c:101-105) -- with segment > 0 the stream of the GPU encoder's history mode */
    const int one_block = block <= 0;
    if (one_block)
        block = n_in > 0 ? n_in : 1;
    if (segment <= 0 || segment > block)
        segment = block;
    if (!one_block && block % segment != 0)
        return LZ77O_E_ARG;

    const int off_bits = lz77o_bitof(SB);
    const int len_bits = lz77o_bitof(LA);
    /* an offset equal to a power-of-two SB does not fit
     * bitof(SB) bits, so the usable window is capped at 2^off_bits - 1. */
    long window = SB;
    if (window > (1L << off_bits) - 1)
        window = (1L << off_bits) - 1;

    memset(out, 0, (size_t)out_cap);
    bitsink sink = { out, out_cap, 0, 0 };
    sink_put(&sink, (uint32_t)SB, 16);
    sink_put(&sink, (uint32_t)LA, 16);

    long tokens = 0;
    for (long base = 0; base < n_in; base += block) {
        const uint8_t *blk = in + base;
        long nb = n_in - base < block ? n_in - base : block;
        long p = 0;
        while (p < nb) {
            /* the greedy parse restarts at every segment: a token never runs
             * past the end of its segment (and so never past its block) */
            long seg_end = (p / segment + 1) * segment;
            if (seg_end > nb)
                seg_end = nb;
            long left = seg_end - p;
            /* lz77.c:87,134 + tree.c:136: len <= min(LA, left) - 1 */
            int max_len = (int)((left < LA ? left : LA) - 1);
            long reach = p < window ? p : window; /* lz77.c:101-105 */
            int m_off, m_len;
            longest_match(blk, p, reach, max_len, &m_off, &m_len);
            sink_put(&sink, (uint32_t)m_off, off_bits);
            sink_put(&sink, (uint32_t)m_len, len_bits);
            sink_put(&sink, blk[p + m_len], 8);
            p += m_len + 1;
            tokens++;
        }
    }
    if (n_tokens)
        *n_tokens = tokens;
    if (sink.overflow)
        return LZ77O_E_SPACE;
    return (sink.bitpos + 7) >> 3;
}

long lz77o_blocked_encode(const uint8_t *in, long n_in, int sb, int la,
                          long block, uint8_t *out, long out_cap,
                          long *n_tokens)
{
    return lz77o_segmented_encode(in, n_in, sb, la, block, 0, out, out_cap,
                                  n_tokens);
}
