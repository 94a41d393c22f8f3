# runs a tool against every library build under variants/ (default: prof_codec on the bench workload)
CMD=${*:-python tools/prof_codec.py text 256 4095 15 3}
for f in variants/*.so; do echo "== $f"; LZ77_B200_LIB=$PWD/$f $CMD 2>&1 | tail -2; done
