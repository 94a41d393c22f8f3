"""Named test inputs shared by the golden generator and the test modules."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from lz77_b200 import synth  # noqa: E402

# (sb, la) pairs of SURVEY.md section 4: both benchmark pairs, a 23-bit token,
# tiny windows, the 9-bit token, and the power-of-two guard.
PARAM_SETS = [(4095, 15), (65535, 255), (1000, 20), (15, 8), (1, 2), (4096, 16)]


def case_input(case: dict) -> bytes:
    kind = case["kind"]
    n = case.get("n", 0)
    seed = case.get("seed", 1234)
    if kind == "literal":
        return case["text"].encode("latin-1")
    if kind == "range256":
        return bytes(range(256)) * (n // 256)
    if kind == "period":
        p = case["period"]
        rng = np.random.default_rng(seed)
        unit = rng.integers(0, 256, p, dtype=np.uint8).tobytes()
        return (unit * (n // p + 1))[:n]
    if kind == "binary2":  # two-symbol alphabet: long hash chains
        rng = np.random.default_rng(seed)
        return (rng.integers(0, 2, n, dtype=np.uint8) + 97).tobytes()
    if kind == "mixed":
        return synth.mixed(n, seed=seed, segment=case.get("segment", 64 << 20)).numpy().tobytes()
    return synth.make(kind, n, seed=seed, device="cpu").numpy().tobytes()


def _c(name, kind, n=0, sb=None, la=None, store=False, **kw):
    d = {"name": name, "kind": kind, "n": n}
    if sb is not None:
        d["sb"] = sb
    if la is not None:
        d["la"] = la
    if store:
        d["store"] = True
    d.update(kw)
    return d


# SURVEY.md Appendix C known-answer vectors first, then reference-produced
# streams the GPU decoder must reproduce (the ones with store=True travel to the
# GPU box as files).
GOLDEN_CASES = [
    _c("kat_empty", "literal", text=""),
    _c("kat_a", "literal", text="a"),
    _c("kat_abc_default", "literal", text="abcabcabcabcX"),
    _c("kat_abc_64k", "literal", text="abcabcabcabcX", sb=65535, la=255),
    _c("kat_abc_t23", "literal", text="abcabcabcabcX", sb=1000, la=20),
    _c("kat_abc_t9", "literal", text="abcabcabcabcX", sb=1, la=2),
    _c("kat_a100", "literal", text="a" * 100),
    _c("kat_zeros_1m", "zeros", n=1 << 20),
    _c("kat_range256x16", "range256", n=4096),
    _c("zeros_64k", "zeros", n=65536, store=True),
    _c("text_256k_default", "zipf_text", n=256 << 10, seed=11, store=True),
    _c("text_300k_64k", "zipf_text", n=300_000, seed=12, sb=65535, la=255, store=True),
    _c("text_100k_t23", "zipf_text", n=100_000, seed=13, sb=1000, la=20, store=True),
    _c("text_20k_s15", "zipf_text", n=20_000, seed=14, sb=15, la=8, store=True),
    _c("text_5k_t9", "zipf_text", n=5_000, seed=15, sb=1, la=2, store=True),
    _c("log_200k_default", "log_like", n=200_000, seed=16, store=True),
    _c("random_100k_default", "random", n=100_000, seed=17, store=True),
    _c("random_200k_64k", "random", n=200_000, seed=18, sb=65535, la=255, store=True),
    _c("period7_50k", "period", n=50_000, period=7, seed=19, store=True),
    _c("period300_80k_64k", "period", n=80_000, period=300, seed=20, sb=65535, la=255, store=True),
    _c("binary2_60k", "binary2", n=60_000, seed=21, store=True),
    _c("text_1m_default", "zipf_text", n=1 << 20, seed=22),
    _c("mixed_768k_l32", "mixed", n=768 << 10, seed=23, sb=8191, la=32, segment=128 << 10),
    _c("pow2_sb4096", "zipf_text", n=64 << 10, seed=24, sb=4096, la=16),
]


# ---- numpy view of a stream's token array (tests only) --------------------------

def stream_params(stream: bytes):
    import math
    sb = stream[0] | stream[1] << 8
    la = stream[2] | stream[3] << 8
    ob = math.ceil(math.log2(sb)) if sb > 1 else 0
    lb = math.ceil(math.log2(la))
    return sb, la, ob, lb, ob + lb + 8


def parse_tokens(stream: bytes):
    """(off, len, lit) arrays of a stream: token k sits at bit 32 + k*T, LSB first
    (lz77.c:246-252, bitio.c:203-239)."""
    import numpy as np
    _, _, ob, lb, T = stream_params(stream)
    body = np.frombuffer(stream, dtype=np.uint8)[4:]
    k = (body.size * 8) // T
    bits = np.unpackbits(body, bitorder="little")[:k * T].reshape(k, T).astype(np.int64)
    w = (bits << np.arange(T, dtype=np.int64)[None, :]).sum(axis=1)
    return w & ((1 << ob) - 1), (w >> ob) & ((1 << lb) - 1), w >> (ob + lb)


def slice_tokens(stream: bytes, a: int, b: int) -> bytes:
    """Standalone stream holding tokens [a, b) of `stream`."""
    import numpy as np
    T = stream_params(stream)[4]
    body = np.frombuffer(stream, dtype=np.uint8)[4:]
    bits = np.unpackbits(body, bitorder="little")[a * T:b * T]
    return stream[:4] + np.packbits(bits, bitorder="little").tobytes()
