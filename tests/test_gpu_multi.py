"""Multi-GPU sharding over NCCL (skipped on a box with fewer than 2 GPUs): the
input is scattered as runs of whole blocks, every rank encodes its shard on its
own GPU, the token payloads are gathered to rank 0 -- the merged stream must be
byte-identical to the single-GPU stream and decode to the input.  The merged
stream is then decoded by all ranks together (even token split, slice sums,
split points nudged to block boundaries) and must give the input again."""
import os
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]


def _nccl_worker(rank, world, port, sb, la, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import lz77_b200
    from lz77_b200 import sharding, synth
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    lz77_b200.init(rank)
    block = lz77_b200.block_size(sb)
    T = lz77_b200.token_bits(sb, la)

    def encode_fn(shard, sb_, la_):
        return lz77_b200.encode_tensor(shard, la=la_, sb=sb_)

    data = synth.zipf_text(n, seed=77, device=dev) if rank == 0 else None
    merged = sharding.encode_sharded(data, n, sb, la, block, T, encode_fn, dev)
    ok = True
    if rank == 0:
        single, _ = lz77_b200.encode_tensor(data, la=la, sb=sb)
        ok = merged.tobytes() == single.cpu().numpy().tobytes()
        ok = ok and lz77_b200.decode(merged.tobytes()) == data.cpu().numpy().tobytes()
    codec = sharding.DecodeCodec(lz77_b200.slice_tokens_tensor, lz77_b200.decode_size_tensor,
                                 lz77_b200.token_at_tensor, lz77_b200.decode_tensor)
    stream = torch.from_numpy(merged.copy()).to(dev) if rank == 0 else None
    out = sharding.decode_sharded(stream, block, T, codec, dev)
    if rank == 0:
        ok = ok and torch.equal(out, data)
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("sb,la", [(4095, 15), (1000, 20), (65535, 255)])
def test_sharded_encode_nccl(sb, la):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000) + sb % 7
    n = 37 * 65536 + 4321 if sb != 65535 else 24 * 524288 + 4321  # >= 1 block of tokens per rank
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, sb, la, n, q))
             for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
