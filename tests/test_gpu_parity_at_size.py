"""Parity at BASELINE.json's sizes through INDEPENDENT decoders / encoders.

A matched encoder/decoder defect (both mishandling, say, a position above 2^32) passes
any self-roundtrip, so at full size the GPU stream is also decoded by

  * the compiled reference, ``oracle/_ref/lz77 -d`` (lz77.c:148-197), on the first, a
    middle and the last 64 MiB of the output -- block-aligned cuts of the token array are
    standalone streams because no match of the block encoder leaves its block -- and
  * the C restatement ``orc.decode`` (seconds even for the whole 256 MiB config),

and the GPU decoder is fed 16 MiB that the compiled reference ENCODED (lz77.c:51-140).
One case decodes to more than 4 GiB (32-bit position arithmetic in the tile decoder).
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SLICE = 64 << 20

# (name, generator, bytes, sb, la) -- the per-GPU shapes of BASELINE.json's configs
CONFIGS = [
    ("configs1_text_256MiB", "zipf_text", 256 << 20, 4095, 15),
    ("configs2_random_1GiB", "random", 1 << 30, 65535, 255),
    ("configs3_share_log_1GiB", "log_like", 1 << 30, 4095, 15),
    ("configs4_share_mixed_2GiB", "mixed", 2 << 30, 65535, 255),
    ("over_4GiB_log_4p25GiB", "log_like", (4 << 30) + (256 << 20), 4095, 15),
]


@pytest.fixture(scope="module")
def lz():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (they never fall back to the CPU)")
    import lz77_b200
    lz77_b200.init(0)
    return lz77_b200


def _ref_decode_file(stream: np.ndarray) -> np.ndarray:
    """oracle/_ref/lz77 -d on tmpfs files (the unmodified reference decoder)."""
    from oracle import ref_binary
    exe = ref_binary()
    assert exe is not None, "oracle/_ref/lz77 did not travel to this box"
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=tmpdir) as d:
        fin, fout = os.path.join(d, "in.lz"), os.path.join(d, "out.bin")
        stream.tofile(fin)
        subprocess.run([str(exe), "-d", "-i", fin, "-o", fout], check=True, timeout=900,
                       capture_output=True)
        return np.fromfile(fout, dtype=np.uint8)


def _slice_stream(lz, stream, lo, hi):
    """Standalone stream of the tokens that decode to output bytes [lo, hi) (block-aligned)."""
    k_lo, p_lo = lz.token_at_tensor(stream, lo)
    k_hi, p_hi = lz.token_at_tensor(stream, hi)
    assert p_lo == lo and p_hi == hi, "a token starts on every block boundary"
    return lz.slice_tokens_tensor(stream, k_lo, k_hi)


@pytest.mark.parametrize("name,kind,n,sb,la", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_gpu_stream_through_reference_decoder_at_size(lz, orc, name, kind, n, sb, la):
    import torch
    from lz77_b200 import synth
    dev = torch.device("cuda", 0)
    src = synth.make(kind, n, seed=1234, device=dev)
    stream, k = lz.encode_tensor(src, la=la, sb=sb)
    T = lz.token_bits(sb, la)
    assert stream.numel() == 4 + (k * T + 7) // 8
    assert lz.decode_size_tensor(stream) == n
    block = lz.block_size(sb)
    # first / middle / last 64 MiB of the output through the compiled reference decoder
    mid = (n // 2) // block * block
    cuts = [(0, min(SLICE, n)), (mid, min(mid + SLICE, n)),
            (max(0, (n - SLICE + block - 1) // block * block), n)]
    from concurrent.futures import ThreadPoolExecutor
    parts = [_slice_stream(lz, stream, lo, hi).cpu().numpy() for lo, hi in cuts]
    with ThreadPoolExecutor(len(cuts)) as ex:   # the threads only wait on the subprocesses
        decoded = list(ex.map(_ref_decode_file, parts))
    for (lo, hi), part, got in zip(cuts, parts, decoded):
        want = src[lo:hi].cpu().numpy()
        assert got.size == want.size and np.array_equal(got, want), (name, lo, hi, "reference -d")
        # the same cut through the GPU decoder as a standalone stream
        back = lz.decode_tensor(torch.from_numpy(np.concatenate([part, np.zeros(16, np.uint8)]))
                                .to(dev)[:part.size])
        assert torch.equal(back, src[lo:hi]), (name, lo, hi, "GPU decode of the cut")
    if n <= (256 << 20):
        # the whole stream through the C restatement of the reference decoder
        assert orc.decode(stream.cpu().numpy().tobytes()) == src.cpu().numpy().tobytes()
    # the whole stream through the GPU decoder (positions above 2^32 in the last config)
    out = lz.decode_tensor(stream)
    assert out.numel() == n and torch.equal(out, src), (name, "GPU decode of the whole stream")


@pytest.mark.parametrize("kind,sb,la", [("zipf_text", 4095, 15), ("random", 65535, 255),
                                        ("log_like", 4095, 15), ("mixed", 65535, 255)])
def test_reference_encoded_slice_through_gpu_decoder(lz, kind, sb, la):
    """16 MiB of each config's generator ENCODED by the compiled reference (unblocked: its
    matches reach back SB bytes from anywhere) and decoded on the GPU."""
    import torch
    from lz77_b200 import synth
    from oracle import ref_run
    n = 16 << 20
    if kind == "mixed":
        data = synth.mixed(n, seed=1234, segment=4 << 20).numpy()
    else:
        data = synth.make(kind, n, seed=1234).numpy()
    ref_stream = ref_run("-c", data.tobytes(), sb=sb, la=la, timeout=900)
    assert lz.decode_size(ref_stream) == n
    assert lz.decode(ref_stream) == data.tobytes()
