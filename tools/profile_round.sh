#!/bin/bash
# The ncu evidence of a round, on one GPU (run under gpurun; outputs under gpurun_out/):
#   $TAG_launches.csv   launch list of `bench.py --steps 2 --warmup 3` (gpu__time_duration per launch)
#   $TAG_bench_full.ncu-rep   ncu --set full of the four main kernels of the timed step of bench.py
#   $TAG_bigwin.ncu-rep, $TAG_jump.ncu-rep   large-window encoder / pointer-jumping decoder
TAG=${1:-r02}
B="python bench.py --no-cli --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    $B --steps 2 --warmup 3 > gpurun_out/${TAG}_launches.log 2>&1
# warm-up 3 steps x (parse, pack, scan, tile) = 12 matching launches, then the timed step
ncu --set full --clock-control none --import-source on \
    -k regex:"lz77_parse_bucket_kernel|lz77_pack_kernel|lz77_decode_scan|lz77_decode_tile_kernel" -s 12 -c 4 \
    -o gpurun_out/${TAG}_bench_full -f $B --steps 1 --warmup 3 > gpurun_out/${TAG}_bench_full.log 2>&1
ncu --set full --clock-control none -k regex:"lz77_block_sort|lz77_parse_bigwin" -s 2 -c 2 \
    -o gpurun_out/${TAG}_bigwin -f python tools/prof_codec.py random 64 65535 255 2 > gpurun_out/${TAG}_bigwin.log 2>&1
ls -la gpurun_out/${TAG}_*
