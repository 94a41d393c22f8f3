#!/usr/bin/env python
"""Decode speed on streams of the reference encoder (unblocked: matches reach
back SB bytes across every tile boundary; decoded by pointer jumping,
decode_jump.cu) next to the block streams of this encoder on the same data.
The stream comes from the byte-identical CPU restatement (slow: ~3 MB/s)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import lz77_b200  # noqa: E402
from lz77_b200 import api, synth  # noqa: E402
from oracle import oracle  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 16
orc = oracle()
lz77_b200.init(0)
for kind, sb, la in (("zipf_text", 4095, 15), ("random", 65535, 255), ("zipf_text", 65535, 255)):
    data = synth.make(kind, mib << 20, seed=5).numpy()
    t0 = time.time()
    cache = Path(f"/tmp/lz77_ref_{kind}_{mib}_{sb}_{la}.npy")  # the CPU encode takes seconds
    if cache.exists():
        ref = np.load(cache)
    else:
        ref = np.frombuffer(orc.ref_encode(data, sb, la), dtype=np.uint8)
        np.save(cache, ref)
    t_cpu = time.time() - t0
    own, _ = api.encode_tensor(torch.from_numpy(data).cuda(), la=la, sb=sb)
    for name, stream in (("reference-style", torch.from_numpy(ref.copy()).cuda()), ("own blocks", own)):
        pad = torch.zeros(((stream.numel() + 15) & ~15) + 16, dtype=torch.uint8, device="cuda")
        pad[:stream.numel()] = stream
        s = pad[:stream.numel()]
        out = torch.empty(((mib << 20) + 31) & ~15, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            back = api.decode_tensor(s, out=out)
            t = api.last_timing()
        assert np.array_equal(back.cpu().numpy(), data)
        print(f"{kind} {mib} MiB sb={sb} la={la} {name:16s}: scan {t['dec_scan_ms']:.3f} ms  tiles "
              f"{t['dec_copy_ms']:.3f} ms  -> {(mib << 20) / (t['dec_copy_ms'] * 1e-3) / 1e9:7.1f} GB/s"
              f"   (cpu encode {t_cpu:.1f} s)")
