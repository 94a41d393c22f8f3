"""lz77_b200 -- Blackwell-native (sm_100a) LZ77 encoder/decoder, drop-in for the
hot path of cstdvd/lz77 (same token bitstream, same encode/decode interface).

  api       ctypes binding of liblz77b200.so (include/lz77_b200.h)
  synth     seeded synthetic inputs of the shapes BASELINE.json names
  sharding  block-range sharding across ranks (one process per GPU)
  csrc/     CUDA kernels, the C ABI and the C command-line program
"""
from .api import (  # noqa: F401
    DEFAULT_LA,
    DEFAULT_SB,
    Lz77Error,
    bitof,
    block_size,
    decode,
    decode_size,
    decode_size_tensor,
    decode_tensor,
    slice_tokens_tensor,
    token_at_tensor,
    encode,
    encode_bound,
    encode_tensor,
    init,
    last_timing,
    load_library,
    set_stream,
    set_timing,
    segment_size,
    shutdown,
    token_bits,
)
