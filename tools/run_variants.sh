# runs tools/prof_codec.py against every library build under variants/
for f in variants/*.so; do echo $f; LZ77_B200_LIB=$PWD/$f python tools/prof_codec.py text 256 4095 15 2; done
