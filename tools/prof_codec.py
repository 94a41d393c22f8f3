#!/usr/bin/env python
"""Small driver for ncu captures: encodes and decodes one synthetic input a few
times through the device entry points (no torch import, so it starts fast).

    ncu ... python tools/prof_codec.py [text|random|zeros] [MiB] [sb] [la] [reps]
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from lz77_b200 import api  # noqa: E402


def text_like(n, seed=1):
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 11)), dtype=np.uint8)) + b" "
             for _ in range(50_000)]
    p = 1.0 / np.arange(1, len(words) + 1) ** 1.1
    p /= p.sum()
    out = bytearray()
    while len(out) < n:
        ids = rng.choice(len(words), size=200_000, p=p)
        out += b"".join(words[i] for i in ids)
    return np.frombuffer(bytes(out[:n]), dtype=np.uint8)


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "text"
    mib = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    sb = int(sys.argv[3]) if len(sys.argv) > 3 else 4095
    la = int(sys.argv[4]) if len(sys.argv) > 4 else 15
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    n = mib << 20
    cache = Path(f"/dev/shm/lz77_prof_{kind}_{mib}.npy")  # variant sweeps reuse the input
    if cache.exists():
        data = np.load(cache)
    else:
        if kind == "text":
            data = text_like(n)
        elif kind == "random":
            data = np.random.default_rng(2).integers(0, 256, n, dtype=np.uint8)
        else:
            data = np.zeros(n, dtype=np.uint8)
        try:
            np.save(cache, data)
        except OSError:
            pass
    api.init(0)
    api.set_host_chunk(0)  # one-shot path, so the per-kernel device times are reported
    cap = api.encode_bound(n, sb, la) + 16
    stream = np.empty(cap, dtype=np.uint8)
    back = np.empty(n + 16, dtype=np.uint8)
    for _ in range(reps):
        c = api.encode_into(data.ctypes.data, n, stream.ctypes.data, cap, la=la, sb=sb)
        t_enc = api.last_timing()
        m = api.decode_into(stream.ctypes.data, c, back.ctypes.data, n)
        t_dec = api.last_timing()
    assert m == n and (back[:n] == data).all()
    print(f"{kind} {mib} MiB sb={sb} la={la}: ratio {n / c:.3f} search {t_enc['enc_search_ms']:.3f} ms "
          f"pack {t_enc['enc_pack_ms']:.3f} ms dscan {t_dec['dec_scan_ms']:.3f} ms "
          f"dcopy {t_dec['dec_copy_ms']:.3f} ms")


if __name__ == "__main__":
    main()
