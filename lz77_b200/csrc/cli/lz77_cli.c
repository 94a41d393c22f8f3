/*
 * lz77_cli.c -- command-line driver with the reference's surface (main.c:59-181):
 *   lz77 -c|-d -i <in> -o <out> [-l 2..255] [-s 0..65535] [-h]
 * Same getopt string plus additive options that leave the file format alone: -g <device>,
 * -G <n> (encode / decode on the first n GPUs: runs of whole blocks per GPU, NCCL
 * scatter / gather inside liblz77b200.so), -p <MiB> (input bytes per library call) and
 * -m <MiB> (decoded bytes per library call), -v (bytes, seconds, GB/s on stderr).  Same limits
 * (main.c:35-38), same messages and exit codes: every usage or open error
 * prints the reference's text on stderr and exits EXIT_FAILURE; -h prints the
 * usage and continues; the last of -c / -d wins; -d ignores -l / -s (the
 * stream header carries them, lz77.c:157-158).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "codec.h"
#include "lz77_b200.h"

enum mode { MODE_NONE, MODE_ENCODE, MODE_DECODE };

static void usage(void)
{
    printf("Usage: lz77 <options>\n");
    printf("  -c : Encode input file to output file.\n");
    printf("  -d : Decode input file to output file.\n");
    printf("  -i <filename> : Name of input file.\n");
    printf("  -o <filename> : Name of output file.\n");
    printf("  -l <value> : Lookahead size (default 15)\n");
    printf("  -s <value> : Search-buffer size (default 4095)\n");
    printf("  -h : Command line options.\n\n");
}

int main(int argc, char *argv[])
{
    enum mode mode = MODE_NONE;
    const char *name_in = NULL, *name_out = NULL;
    int la = -1, sb = -1; /* -1: library default (lz77.c:65-66) */
    FILE *plain = NULL;
    struct bitFILE *packed = NULL;
    int opt;

    while ((opt = getopt(argc, argv, "cdi:o:l:s:hg:G:p:m:v")) != -1) {
        switch (opt) {
        case 'c':
            mode = MODE_ENCODE;
            break;
        case 'd':
            mode = MODE_DECODE;
            break;
        case 'i':
            if (name_in != NULL) {
                fprintf(stderr, "Multiple input files not allowed.\n");
                goto fail;
            }
            name_in = optarg;
            break;
        case 'o':
            if (name_out != NULL) {
                fprintf(stderr, "Multiple output files not allowed.\n");
                goto fail;
            }
            name_out = optarg;
            break;
        case 'l':
            la = atoi(optarg);
            if (la < LZ77_MIN_LA || la > LZ77_MAX_LA) {
                fprintf(stderr, "Bad lookahead size value.\n");
                goto fail;
            }
            break;
        case 's':
            sb = atoi(optarg);
            if (sb < LZ77_MIN_SB || sb > LZ77_MAX_SB) {
                fprintf(stderr, "Bad search-buffer size value.\n");
                goto fail;
            }
            break;
        case 'h':
            usage();
            break;
        case 'g':
            lz77_cli_set_device(atoi(optarg));
            break;
        case 'G':
            lz77_cli_set_gpus(atoi(optarg));
            break;
        case 'p':
            lz77_cli_set_piece_mib(atol(optarg));
            break;
        case 'm':
            lz77_cli_set_out_mib(atol(optarg));
            break;
        case 'v':
            lz77_cli_set_verbose(1);
            break;
        default:
            break;
        }
    }

    if (name_in == NULL) {
        fprintf(stderr, "Input file must be provided\n");
        goto fail;
    }
    if (name_out == NULL) {
        fprintf(stderr, "Output file must be provided\n");
        goto fail;
    }

    if (mode == MODE_ENCODE) {
        if ((plain = fopen(name_in, "rb")) == NULL) {
            perror("Opening input file");
            goto fail;
        }
        if ((packed = bitIO_open(name_out, BIT_IO_W)) == NULL) {
            perror("Opening output file");
            goto fail;
        }
        encode(plain, packed, la, sb);
    } else if (mode == MODE_DECODE) {
        if ((packed = bitIO_open(name_in, BIT_IO_R)) == NULL) {
            perror("Opening input file");
            goto fail;
        }
        if ((plain = fopen(name_out, "wb")) == NULL) {
            perror("Opening output file");
            goto fail;
        }
        decode(packed, plain);
    } else {
        fprintf(stderr, "Select ENCODE or DECODE mode\n");
        goto fail;
    }

    fclose(plain);
    bitIO_close(packed);
    return 0;

fail:
    if (plain != NULL)
        fclose(plain);
    if (packed != NULL)
        bitIO_close(packed);
    return EXIT_FAILURE;
}
