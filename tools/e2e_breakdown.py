#!/usr/bin/env python
"""Wall-clock breakdown of the host entry points (pinned buffers) on the bench
workload: encode and decode separately, for a few host chunk sizes."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import lz77_b200  # noqa: E402
from lz77_b200 import api, synth  # noqa: E402

n = 256 << 20
lz77_b200.init(0)
src = synth.zipf_text(n, seed=1234, device="cuda")
cap = api.encode_bound(n) + 16
h_in, h_stream, h_out = api.PinnedBuffer(n), api.PinnedBuffer(cap), api.PinnedBuffer(n + 16)
h_in.array[:] = src.cpu().numpy()
for chunk_mib in (0, 4, 8, 12, 16, 24, 32):
    api.set_host_chunk(chunk_mib << 20)
    for _ in range(2):
        c = api.encode_into(h_in.ptr, n, h_stream.ptr, cap)
        m = api.decode_into(h_stream.ptr, c, h_out.ptr, n)
    torch.cuda.synchronize()
    te = td = 0.0
    reps = 5
    for _ in range(reps):
        t0 = time.perf_counter()
        c = api.encode_into(h_in.ptr, n, h_stream.ptr, cap)
        t1 = time.perf_counter()
        span = api.last_timing()["enc_search_ms"]
        m = api.decode_into(h_stream.ptr, c, h_out.ptr, n)
        t2 = time.perf_counter()
        te += t1 - t0
        td += t2 - t1
    assert m == n and (h_out.array[:n] == h_in.array).all()
    print(f"chunk {chunk_mib:3d} MiB: encode {te / reps * 1e3:6.2f} ms (compute-stream span {span:6.2f} ms, "
          f"{n / (te / reps) / 1e9:5.1f} GB/s)  "
          f"decode {td / reps * 1e3:6.2f} ms ({n / (td / reps) / 1e9:5.1f} GB/s)")
