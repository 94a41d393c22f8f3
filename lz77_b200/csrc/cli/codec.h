/*
 * codec.h -- the two entry points the command-line driver calls, with the
 * reference's signatures (lz77.h:14-15) and its bit-file handle (bitio.h:18,
 * 25-27), implemented on top of the C ABI in include/lz77_b200.h.
 */
#ifndef LZ77_CLI_CODEC_H
#define LZ77_CLI_CODEC_H

#include <stdio.h>

#define BIT_IO_W 0 /* bitio.h:12 */
#define BIT_IO_R 1 /* bitio.h:13 */

struct bitFILE; /* opaque, as in bitio.h:18 */

struct bitFILE *bitIO_open(const char *path, int mode); /* bitio.c:124 */
int bitIO_close(struct bitFILE *bitF);                  /* bitio.c:171 */

/* la / sb of -1 select the defaults (lz77.c:65-66) */
void encode(FILE *file, struct bitFILE *out, int la, int sb); /* lz77.c:51  */
void decode(struct bitFILE *file, FILE *out);                 /* lz77.c:148 */

/* additive knobs: CUDA device the codec binds to (default 0); number of GPUs (default 1:
 * > 1 shards every piece over the first n devices, lz77_mgpu_*); input bytes per library
 * call (encode: default 1 GiB; decode: a quarter of it in stream bytes) and decoded bytes
 * per library call (default 4 GiB) -- files of any size stream through in pieces */
void lz77_cli_set_device(int device);
void lz77_cli_set_gpus(int n);
void lz77_cli_set_piece_mib(long mib);
void lz77_cli_set_out_mib(long mib);
void lz77_cli_set_verbose(int on); /* -v: bytes, seconds and GB/s on stderr */

#endif
