#!/usr/bin/env python
"""Soak test: repeats device encode/decode many times on varied inputs and checks
every result (the decode kernel synchronises through flags and spin waits, so a
race would show up as a rare mismatch or a hang -- run under `timeout`)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import lz77_b200  # noqa: E402
from lz77_b200 import api, synth  # noqa: E402
from oracle import oracle  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
lz77_b200.init(0)
orc = oracle()
rng = np.random.default_rng(7)
t0 = time.time()
n_checks = 0
for kind in ("zipf_text", "random", "log_like", "zeros", "mixed"):
    for sb, la in ((4095, 15), (65535, 255), (1000, 20), (255, 255)):
        n = int(rng.integers(1 << 20, 24 << 20))
        src = synth.mixed(n, seed=3, device="cuda", segment=1 << 20) if kind == "mixed" \
            else synth.make(kind, n, seed=int(rng.integers(1 << 30)), device="cuda")
        ref_stream = None
        first = None
        for r in range(reps):
            s, k = api.encode_tensor(src, la=la, sb=sb)
            if first is None:
                first = s.clone()
            else:
                assert torch.equal(s, first), ("encode not deterministic", kind, sb, la, r)
            back = api.decode_tensor(s)
            assert torch.equal(back, src), ("decode mismatch", kind, sb, la, r)
            n_checks += 1
        # reference-style stream (cross-tile dependencies), a few repetitions
        small = src[:600_000].cpu().numpy()
        ref = torch.from_numpy(np.frombuffer(orc.ref_encode(small, sb, la), dtype=np.uint8).copy())
        pad = torch.zeros(((ref.numel() + 15) & ~15) + 16, dtype=torch.uint8, device="cuda")
        pad[:ref.numel()] = ref.cuda()
        for r in range(10):
            back = api.decode_tensor(pad[:ref.numel()])
            assert np.array_equal(back.cpu().numpy(), small), ("ref-style mismatch", kind, sb, la, r)
            n_checks += 1
        print(f"{kind:10s} sb={sb:5d} la={la:3d} n={n:9d} ok  ({time.time() - t0:.0f} s)", flush=True)
print(f"soak ok: {n_checks} roundtrips verified")
