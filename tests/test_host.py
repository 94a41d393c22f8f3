"""CPU-side tests: the C-ABI library loads and exports what include/lz77_b200.h
declares, format arithmetic, the no-GPU error path, the command-line surface,
and block-range sharding (including a world_size-2 gloo run).  The oracle
stands in for the GPU codec only here, as the sharding module takes the codec
as a callable."""
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from _cases import PARAM_SETS

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "lz77_b200.h"
CLI = ROOT / "lz77_b200" / "bin" / "lz77"


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not (ROOT / "lz77_b200" / "liblz77b200.so").exists() or not CLI.exists():
        g.build()
    from lz77_b200 import api
    return api.load_library()


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(lz77_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from lz77_b200 import api
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/lz77_b200.h but not exported"
    assert set(api.EXPORTS) == set(names)


def test_format_arithmetic_matches_oracle(lib, orc):
    for n in list(range(1, 70)) + [255, 256, 257, 4095, 4096, 65535]:
        assert lib.lz77_bitof(n) == orc.bitof(n)
    for sb, la in PARAM_SETS:
        assert lib.lz77_token_bits(sb, la) == orc.token_bits(sb, la)
        for n in (0, 1, 1000, 1 << 20):
            assert lib.lz77_gpu_encode_bound(n, sb, la) == orc.encode_bound(n, sb, la)
    assert lib.lz77_gpu_encode_bound(10, -1, -1) == 4 + 30
    assert lib.lz77_gpu_block_size(4095) == 65536
    assert lib.lz77_gpu_block_size(65535) == 524288
    for sb, la in PARAM_SETS:
        seg = lib.lz77_gpu_segment_size(sb, la)
        assert seg in (256, 512, 1024) and 65536 % seg == 0
    assert lib.lz77_gpu_segment_size(65535, 255) == 1024


def test_no_device_fails_loudly(lib):
    """Without a GPU every compute entry point reports an error -- there is no
    CPU fallback inside the product."""
    if lib.lz77_gpu_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    import ctypes as C
    from lz77_b200 import api
    assert lib.lz77_gpu_init(0) == api.E_NODEVICE
    buf = (C.c_uint8 * 64)()
    n = C.c_long(0)
    assert lib.lz77_gpu_encode(C.addressof(buf), 16, -1, -1, C.addressof(buf), 64,
                               C.byref(n)) == api.E_NODEVICE
    assert lib.lz77_gpu_decode(C.addressof(buf), 16, C.addressof(buf), 64,
                               C.byref(n)) == api.E_NODEVICE
    with pytest.raises(api.Lz77Error):
        api.encode(b"abc")
    with pytest.raises(api.Lz77Error):
        api.decode(b"\xff\x0f\x0f\x00")


def _run(args):
    return subprocess.run([str(CLI), *args], capture_output=True, text=True)


def test_cli_surface_matches_reference(lib, tmp_path):
    """Same options, messages and exit codes as reference main.c:69-139,173-180."""
    f = tmp_path / "in.bin"
    f.write_bytes(b"hello")
    o = str(tmp_path / "out")
    cases = [
        (["-c", "-o", o], "Input file must be provided"),
        (["-c", "-i", str(f)], "Output file must be provided"),
        (["-i", str(f), "-o", o], "Select ENCODE or DECODE mode"),
        (["-c", "-i", str(f), "-i", str(f), "-o", o], "Multiple input files not allowed."),
        (["-c", "-i", str(f), "-o", o, "-o", o], "Multiple output files not allowed."),
        (["-c", "-i", str(f), "-o", o, "-l", "1"], "Bad lookahead size value."),
        (["-c", "-i", str(f), "-o", o, "-l", "256"], "Bad lookahead size value."),
        (["-c", "-i", str(f), "-o", o, "-s", "65536"], "Bad search-buffer size value."),
        (["-c", "-i", str(tmp_path / "missing"), "-o", o], "Opening input file"),
    ]
    for args, msg in cases:
        r = _run(args)
        assert r.returncode == 1, args
        assert msg in r.stderr, (args, r.stderr)
    r = _run(["-h"])
    assert r.returncode == 1 and "Usage: lz77 <options>" in r.stdout
    assert "Lookahead size (default 15)" in r.stdout
    assert "Input file must be provided" in r.stderr     # -h does not exit by itself


def test_cli_matches_reference_binary_messages(lib, tmp_path, ref_available):
    if not ref_available:
        pytest.skip("oracle/_ref/lz77 not built")
    from oracle import ref_binary
    f = tmp_path / "in.bin"
    f.write_bytes(b"hello")
    o = str(tmp_path / "out")
    for args in (["-c", "-o", o], ["-i", str(f), "-o", o], ["-c", "-i", str(f), "-o", o, "-l", "1"],
                 ["-h"]):
        ours = _run(args)
        ref = subprocess.run([str(ref_binary()), *args], capture_output=True, text=True)
        assert ours.returncode == ref.returncode
        assert ours.stderr == ref.stderr
        assert ours.stdout == ref.stdout


# ---- sharding ---------------------------------------------------------------

def test_shard_ranges_cover_whole_blocks():
    from lz77_b200.sharding import shard_ranges
    for n in (0, 1, 65535, 65536, 65537, 10 * 65536 + 5, 1 << 24):
        for world in (1, 2, 3, 4, 8):
            rs = shard_ranges(n, world, 65536)
            assert len(rs) == world and rs[0][0] == 0 and rs[-1][1] == n
            for (a, b), (c, d) in zip(rs, rs[1:]):
                assert b == c and a <= b
            for a, b in rs:
                assert a % 65536 == 0 or a == n
            sizes = [(b - a + 65535) // 65536 for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("sb,la", [(4095, 15), (65535, 255), (1000, 20), (1, 2), (15, 8)])
def test_merged_shards_equal_single_stream(orc, lib, sb, la):
    """Shards encoded independently and merged == the whole input encoded at
    once, byte for byte, also when tokens are not byte aligned (23- and 9-bit)."""
    from lz77_b200 import synth
    from lz77_b200.sharding import merge_payloads, shard_ranges, split_stream
    block = lib.lz77_gpu_block_size(sb)
    seg = lib.lz77_gpu_segment_size(sb, la)
    T = lib.lz77_token_bits(sb, la)
    data = synth.zipf_text(5 * block + 12_345, seed=4).numpy()
    whole, k_whole = orc.blocked_encode(data, sb, la, block, seg)
    for world in (2, 3, 8):
        payloads, counts, header = [], [], None
        for lo, hi in shard_ranges(data.size, world, block):
            s, k = orc.blocked_encode(data[lo:hi], sb, la, block, seg)
            h, p = split_stream(np.frombuffer(s, dtype=np.uint8))
            header = header or h
            payloads.append(p)
            counts.append(k)
        assert sum(counts) == k_whole
        merged = merge_payloads(header, payloads, counts, T)
        assert merged.tobytes() == whole


def _gloo_worker(rank, world, port, sb, la, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    from lz77_b200 import api, sharding, synth
    from oracle import oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle()
    lib = api.load_library()
    block, seg = lib.lz77_gpu_block_size(sb), lib.lz77_gpu_segment_size(sb, la)
    T = lib.lz77_token_bits(sb, la)

    def encode_fn(shard, sb_, la_):  # stand-in for lz77_b200.encode_tensor (tests only)
        s, k = orc.blocked_encode(shard.numpy(), sb_, la_, block, seg)
        return torch.frombuffer(bytearray(s), dtype=torch.uint8), k

    data = synth.zipf_text(n, seed=8) if rank == 0 else None
    merged = sharding.encode_sharded(data, n, sb, la, block, T, encode_fn, "cpu")
    if rank == 0:
        whole, _ = orc.blocked_encode(data.numpy(), sb, la, block, seg)
        q.put(merged.tobytes() == whole and orc.decode(merged.tobytes()) == data.numpy().tobytes())
    dist.barrier()
    dist.destroy_process_group()


def _gloo_decode_worker(rank, world, port, sb, la, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    from _cases import parse_tokens, slice_tokens
    from lz77_b200 import api, sharding, synth
    from oracle import oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle()
    lib = api.load_library()
    block, seg = lib.lz77_gpu_block_size(sb), lib.lz77_gpu_segment_size(sb, la)
    T = lib.lz77_token_bits(sb, la)

    def t2b(t):
        return t.numpy().tobytes()

    def b2t(b):
        return torch.frombuffer(bytearray(b), dtype=torch.uint8)

    def token_at(s, pos):  # numpy stand-in for lz77_gpu_token_at_device
        _, length, _ = parse_tokens(t2b(s))
        start = np.concatenate([[0], np.cumsum(length + 1)])
        k = int(np.searchsorted(start, pos, side="right") - 1)
        return k, int(start[k])

    codec = sharding.DecodeCodec(
        slice_tokens=lambda s, a, b: b2t(slice_tokens(t2b(s), a, b)),
        decode_size=lambda s: int((parse_tokens(t2b(s))[1] + 1).sum()),
        token_at=token_at,
        decode=lambda s: b2t(orc.decode(t2b(s))))
    data = synth.zipf_text(abs(n), seed=8).numpy()
    stream = None
    if rank == 0:
        if n < 0:  # what the reference encoder writes: no token on the block boundaries
            whole = orc.ref_encode(data, sb, la)
        else:
            whole, _ = orc.blocked_encode(data, sb, la, block, seg)
        stream = b2t(whole)
    if n < 0:
        # every rank must raise (none may be left waiting in a collective)
        try:
            sharding.decode_sharded(stream, block, T, codec, "cpu")
            q.put((rank, "no error"))
        except ValueError as e:
            q.put((rank, str(e)))
        dist.barrier()
        dist.destroy_process_group()
        return
    out = sharding.decode_sharded(stream, block, T, codec, "cpu")
    if rank == 0:
        k = ((stream.numel() - 4) * 8) // T
        q.put((t2b(out) == data.tobytes(), k // world >= block))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("sb,la,n,split", [(4095, 15, 24 * 65536 + 999, True),
                                           (1000, 20, 16 * 65536 + 5, True),
                                           (4095, 15, 70_000, False)])
def test_sharded_decode_world2_gloo(sb, la, n, split):
    """One block-structured stream decoded by a 2-rank gloo group: even token split,
    all-gather of the slice sums, split points nudged to block boundaries (the last
    case is too small to split and decodes on root alone)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + (n % 13)
    procs = [ctx.Process(target=_gloo_decode_worker, args=(r, 2, port, sb, la, n, q))
             for r in range(2)]
    for p in procs:
        p.start()
    ok, was_split = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and was_split == split


def test_sharded_decode_reference_stream_fails_on_every_rank_gloo():
    """A stream of the reference encoder cannot shard: both ranks raise ValueError after
    the all-gather of the split points; neither is left waiting in a collective."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_decode_worker, args=(r, 2, port, 4095, 15, -700_001, q))
             for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert set(got) == {0, 1}
    assert all("not a stream of the block encoder" in msg for msg in got.values()), got


def test_shard_range_matches_the_python_bookkeeping():
    """lz77_shard_range (what the library's NCCL path uses) == sharding.shard_ranges."""
    from lz77_b200 import api
    from lz77_b200.sharding import shard_ranges
    for n, world, block in [(0, 2, 65536), (1, 3, 65536), (5 * 65536 + 17, 2, 65536),
                            (37 * 65536 + 4321, 4, 65536), (24 * 524288 + 4321, 8, 524288),
                            (65536, 8, 65536), (1 << 32, 4, 65536)]:
        want = shard_ranges(n, world, block)
        got = [api.shard_range(n, world, block, r) for r in range(world)]
        assert got == want, (n, world, block)


@pytest.mark.parametrize("sb,la", [(4095, 15), (1000, 20)])
def test_sharded_encode_world2_gloo(sb, la):
    """scatter -> per-rank encode -> gather over a 2-rank gloo group gives the
    single-GPU stream."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if sb == 1000 else 0)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, sb, la, 3 * 65536 + 999, q))
             for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


# --------------------------------------------------------------------------- #
# the command-line program's host logic on CPUs (no GPU): the unmodified CLI sources
# linked against tests/cli_stand_in/stand_in.c, a stand-in for the library built from
# the oracle (test infrastructure; the product never links it)
# --------------------------------------------------------------------------- #

@pytest.fixture(scope="module")
def cli_on_cpu(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli_stand_in")
    inc = ["-I", str(ROOT / "include"), "-I", str(ROOT / "oracle")]
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", *inc, "-o", str(d / "liblz77b200.so"),
                    str(ROOT / "tests" / "cli_stand_in" / "stand_in.c"),
                    str(ROOT / "oracle" / "lz77_oracle.c"), "-lm"], check=True)
    cli_src = ROOT / "lz77_b200" / "csrc" / "cli"
    subprocess.run(["gcc", "-O2", "-std=gnu11", "-I", str(ROOT / "include"), "-o", str(d / "lz77"),
                    str(cli_src / "lz77_cli.c"), str(cli_src / "codec.c"), "-L", str(d),
                    "-llz77b200", "-lpthread", "-Wl,-rpath," + str(d)], check=True)
    return d / "lz77"


@pytest.mark.parametrize("sb,la", [(4095, 15), (1000, 20), (65535, 255)])
def test_cli_piece_logic_on_cpu(cli_on_cpu, orc, tmp_path, sb, la):
    """codec.c without a GPU: the encoder's pieces give the stream of a single call; the
    decoder's pieces (whole tokens, the output tail carried as literal tokens -- assembled
    in place for byte-aligned tokens, bit by bit for the 23-bit token --, read-ahead and
    writer threads, the smaller piece on LZ77_E_SPACE) reproduce the input from streams of
    the block encoder and of the reference encoder (matches across every piece seam)."""
    from lz77_b200 import synth
    data = synth.zipf_text(1_300_001, seed=21).numpy().tobytes() + bytes(700_000)
    fin, one, pieces, back = (tmp_path / n for n in ("in.bin", "one.lz", "pieces.lz", "back.bin"))
    fin.write_bytes(data)
    args = ["-s", str(sb), "-l", str(la)]
    subprocess.run([str(cli_on_cpu), "-c", "-i", str(fin), "-o", str(one), *args], check=True)
    subprocess.run([str(cli_on_cpu), "-c", "-i", str(fin), "-o", str(pieces), "-p", "1", *args],
                   check=True)
    assert pieces.read_bytes() == one.read_bytes()
    assert orc.decode(one.read_bytes()) == data
    ref = tmp_path / "ref.lz"
    ref.write_bytes(orc.ref_encode(data[:900_000], sb, la))
    for stream, expect in ((one, data), (ref, data[:900_000])):
        for extra in ([], ["-p", "1"], ["-p", "1", "-m", "1"]):
            back.unlink(missing_ok=True)
            subprocess.run([str(cli_on_cpu), "-d", "-i", str(stream), "-o", str(back), *extra],
                           check=True)
            assert back.read_bytes() == expect, (stream.name, extra)


def test_cli_edge_files_on_cpu(cli_on_cpu, tmp_path):
    """Empty and tiny files, and a stream cut in the middle of a token (lz77.c:271-280: the
    trailing bits are padding) through the piece logic."""
    for name, data in (("empty", b""), ("one", b"a"), ("abc", b"abcabcabcabcX")):
        fin, enc, back = tmp_path / f"{name}.bin", tmp_path / f"{name}.lz", tmp_path / f"{name}.out"
        fin.write_bytes(data)
        subprocess.run([str(cli_on_cpu), "-c", "-i", str(fin), "-o", str(enc)], check=True)
        subprocess.run([str(cli_on_cpu), "-d", "-i", str(enc), "-o", str(back)], check=True)
        assert back.read_bytes() == data
    cut = tmp_path / "cut.lz"
    cut.write_bytes((tmp_path / "abc.lz").read_bytes()[:-1])
    subprocess.run([str(cli_on_cpu), "-d", "-i", str(cut), "-o", str(tmp_path / "cut.out")], check=True)
    assert b"abcabcabcabcX".startswith((tmp_path / "cut.out").read_bytes())
