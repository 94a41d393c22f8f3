#!/usr/bin/env python
"""Small inputs through the encode kernels only, for

    compute-sanitizer --tool racecheck python tools/racecheck_encode.py

(the decode tile kernel orders its shared-memory traffic with a ready bitmap instead of
barriers and is reported by racecheck by design; the encoders use
barriers only and must come out clean)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import lz77_b200 as lz  # noqa: E402
from lz77_b200 import synth  # noqa: E402

lz.init(0)
for kind, n, sb, la in (("zipf_text", 300_001, 4095, 15), ("random", 100_000, 1000, 20),
                        ("zeros", 70_000, 4095, 15), ("zipf_text", 700_000, 65535, 255),
                        ("zipf_text", 50_000, 1, 15)):
    src = synth.make(kind, n, seed=3, device="cuda")
    stream, ntok = lz.encode_tensor(src, la=la, sb=sb)
    torch.cuda.synchronize()
    print(kind, n, sb, la, "->", stream.numel(), "bytes,", ntok, "tokens")
print("racecheck_encode: done")
