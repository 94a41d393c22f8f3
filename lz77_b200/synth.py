"""Seeded synthetic inputs of the shapes BASELINE.json names.

Everything is generated with torch so the same code runs on the CPU (tests, the
CPU-baseline sample) and on the GPU (full-size bench inputs that never touch the
host).  The byte streams are deterministic per (seed, device type).

  zeros      config 1: all-zero file
  zipf_text  config 2: 50,000 pseudo-words of 2-10 lowercase letters, word rank
             probability ~ 1/r^1.1, joined by single spaces
  random     config 3: uniform random bytes (incompressible)
  log_like   config 4: one 4096-byte text line template repeated, bytes 0-9 of
             every line overwritten by a zero-padded decimal line counter
  mixed      config 5: fixed-size segments cycling text / log / random
"""
from __future__ import annotations

import torch

VOCAB_WORDS = 50_000
ZIPF_EXPONENT = 1.1
LOG_PERIOD = 4096


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def zeros(n: int, device="cpu") -> torch.Tensor:
    return torch.zeros(int(n), dtype=torch.uint8, device=device)


def random_bytes(n: int, seed: int = 1234, device="cpu") -> torch.Tensor:
    g = _gen(seed, device)
    out = torch.empty(int(n), dtype=torch.uint8, device=device)
    step = 1 << 28
    for s in range(0, int(n), step):
        e = min(int(n), s + step)
        out[s:e] = torch.randint(0, 256, (e - s,), generator=g, device=device,
                                 dtype=torch.int32).to(torch.uint8)
    return out


def _vocabulary(seed: int, device):
    """Flat byte table of the vocabulary, each word followed by one space."""
    g = _gen(seed ^ 0x5EED, device)
    lens = torch.randint(2, 11, (VOCAB_WORDS,), generator=g, device=device)
    width = lens + 1  # trailing space
    start = torch.cumsum(width, 0) - width
    total = int(width.sum())
    flat = torch.randint(97, 123, (total,), generator=g, device=device,
                         dtype=torch.int32).to(torch.uint8)
    flat[start + lens] = 32
    ranks = torch.arange(1, VOCAB_WORDS + 1, device=device, dtype=torch.float64)
    p = ranks.pow(-ZIPF_EXPONENT)
    cdf = torch.cumsum(p / p.sum(), 0)
    return flat, start, width, cdf


def zipf_text(n: int, seed: int = 1234, device="cpu") -> torch.Tensor:
    n = int(n)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    if n == 0:
        return out
    flat, start, width, cdf = _vocabulary(seed, device)
    g = _gen(seed, device)
    words_per_chunk = 1 << 21
    filled = 0
    while filled < n:
        u = torch.rand(words_per_chunk, generator=g, device=device, dtype=torch.float64)
        ids = torch.searchsorted(cdf, u).clamp_(max=VOCAB_WORDS - 1)
        w = width[ids]
        ends = torch.cumsum(w, 0)
        total = int(ends[-1])
        # byte j of the chunk belongs to word k: source = flat[start[k] + j - begin[k]]
        shift = start[ids] - (ends - w)
        src = torch.repeat_interleave(shift, w, output_size=total)
        src += torch.arange(total, device=device)
        take = min(total, n - filled)
        out[filled:filled + take] = flat[src[:take]]
        filled += take
    return out


def log_like(n: int, seed: int = 1234, device="cpu", period: int = LOG_PERIOD) -> torch.Tensor:
    n = int(n)
    if n == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    template = zipf_text(period, seed=seed + 7, device=device)
    lines = (n + period - 1) // period
    out = template.repeat(lines)
    view = out.view(lines, period)
    counter = torch.arange(lines, device=device, dtype=torch.int64)
    for d in range(10):
        digit = (counter // (10 ** (9 - d))) % 10 + 48
        view[:, d] = digit.to(torch.uint8)
    return out[:n].contiguous()


def mixed(n: int, seed: int = 1234, device="cpu", segment: int = 64 << 20) -> torch.Tensor:
    n = int(n)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    kinds = (zipf_text, log_like, random_bytes)
    pos, idx = 0, 0
    while pos < n:
        take = min(segment, n - pos)
        out[pos:pos + take] = kinds[idx % 3](take, seed=seed + 101 * idx, device=device)
        pos += take
        idx += 1
    return out


GENERATORS = {
    "zeros": lambda n, seed=0, device="cpu": zeros(n, device=device),
    "zipf_text": zipf_text,
    "random": random_bytes,
    "log_like": log_like,
    "mixed": mixed,
}


def make(kind: str, n: int, seed: int = 1234, device="cpu") -> torch.Tensor:
    """Generate ``n`` bytes of the named workload as a uint8 tensor."""
    try:
        fn = GENERATORS[kind]
    except KeyError as exc:
        raise ValueError(f"unknown synthetic workload {kind!r}") from exc
    return fn(n, seed=seed, device=device)
