#!/usr/bin/env python
"""Small inputs through every kernel, for compute-sanitizer:

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py

Covers both encoders (window <= 8191 with the backward bucket walk, and 64 KiB), the fused
search + pack path, history mode, the tile decoder and both token scans, the
pointer-jumping decoder (an unblocked stream of random tokens), the chunked host
pipelines and the token-array helpers; every result is checked on the way."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import lz77_b200 as lz  # noqa: E402
from lz77_b200 import api, synth  # noqa: E402
from test_gpu_parity import _pack_tokens, _random_tokens  # noqa: E402
from oracle import oracle  # noqa: E402

orc = oracle()
lz.init(0)
for kind, n, sb, la in (("zipf_text", 700_001, 4095, 15), ("random", 300_000, 1000, 20),
                        ("zipf_text", 1_300_000, 65535, 255), ("zeros", 200_000, 65535, 255)):
    data = synth.make(kind, n, seed=3).numpy().tobytes()
    enc = lz.encode(data, la=la, sb=sb)
    assert orc.decode(enc) == data and lz.decode(enc) == data, (kind, sb, la)
    s = torch.frombuffer(bytearray(enc), dtype=torch.uint8).cuda()
    k = ((len(enc) - 4) * 8) // lz.token_bits(sb, la)
    sub = lz.slice_tokens_tensor(s, k // 3, k)
    assert lz.decode_size_tensor(s) == n and sub.numel() > 4
    tok, pos = lz.token_at_tensor(s, n // 2)
    assert 0 <= tok < k and pos <= n // 2
rng = np.random.default_rng(1)
for sb, la, k in ((4095, 15, 300_000), (65535, 255, 30_000), (1000, 20, 100_000)):
    stream = _pack_tokens(*_random_tokens(rng, k, sb, la), sb, la)
    assert lz.decode(stream) == orc.decode(stream), ("unblocked", sb, la)
# encoder options: fused pack (look-back, spill rows, parked tiles), history mode
for fused, hist in ((True, False), (False, True), (True, True)):
    api.set_fused_pack(fused)
    api.set_history(hist)
    for kind, n, sb, la in (("zipf_text", 400_001, 4095, 15), ("random", 150_000, 4095, 15),
                            ("zeros", 100_000, 4095, 15), ("zipf_text", 1_200_000, 65535, 255)):
        data = synth.make(kind, n, seed=5).numpy().tobytes()
        enc = lz.encode(data, la=la, sb=sb)
        assert orc.decode(enc) == data and lz.decode(enc) == data, (fused, hist, kind, sb, la)
api.set_fused_pack(False)
api.set_history(False)
api.set_host_chunk(1 << 20)
data = synth.zipf_text(5_000_000, seed=4).numpy().tobytes()
enc = lz.encode(data)
assert lz.decode(enc) == data
api.set_fused_pack(True)   # launches of one call on two streams, look-back across them
assert lz.decode(lz.encode(data)) == data
api.set_fused_pack(False)
stream = _pack_tokens(*_random_tokens(rng, 1_200_000, 4095, 15), 4095, 15)
assert lz.decode(stream) == orc.decode(stream)
print("sanitize_smoke: ok")
