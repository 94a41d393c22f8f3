"""One input, one stream, several GPUs -- the library's own NCCL path (comm.cu) behind
the C ABI (skipped on a box with fewer than 2 GPUs).

  * ranks = processes: root scatters runs of whole blocks over NCCL, every rank encodes
    its run, the payloads are gathered into ONE stream that must be byte-identical to the
    single-GPU stream (one header, lz77.c:74-75; contiguous fixed-width tokens,
    lz77.c:246-252) and decode through the oracle; the merged stream is then decoded by
    all ranks together and must give the input again;
  * a stream of the reference encoder cannot shard: LZ77_E_STREAM on EVERY rank, no hang;
  * ranks = threads of one process (lz77_mgpu_*, what the command-line -G uses).
"""
import os
import subprocess
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
    from lz77_b200 import api, synth
    from oracle import oracle
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    api.comm_init_torch(dev)   # the id travels through torch; the communicator is the library's
    ok, why = True, ""

    # --- encode: merged stream == single-GPU stream ------------------------------
    data = synth.zipf_text(n, seed=77, device=dev) if rank == 0 else None
    merged, k_total = api.encode_sharded_tensor(data, la=la, sb=sb)
    stats_enc = api.comm_stats()
    if rank == 0:
        single, k_single = lz77_b200.encode_tensor(data, la=la, sb=sb)
        if not (torch.equal(merged, single) and k_total == k_single):
            ok, why = False, "merged stream differs from the single-GPU stream"
        raw = data.cpu().numpy().tobytes()
        if ok and oracle().decode(merged.cpu().numpy().tobytes()) != raw:
            ok, why = False, "oracle decode of the merged stream differs"
        if ok and not (stats_enc["sent_bytes"] > 0 and stats_enc["recv_bytes"] > 0):
            ok, why = False, f"no bytes moved through NCCL: {stats_enc}"
    # --- decode: all ranks together -------------------------------------------------
    stream = merged.clone() if rank == 0 else None
    out = api.decode_sharded_tensor(stream, out_cap=n)
    if rank == 0 and ok and not torch.equal(out, data):
        ok, why = False, "sharded decode differs from the input"
    # --- too small to split: root decodes alone, every rank returns -------------------
    small = synth.zipf_text(70_000, seed=5, device=dev) if rank == 0 else None
    s_small, _ = api.encode_sharded_tensor(small, la=la, sb=sb)
    o_small = api.decode_sharded_tensor(s_small.clone() if rank == 0 else None, out_cap=70_000)
    if rank == 0 and ok and not torch.equal(o_small, small):
        ok, why = False, "small input roundtrip differs"
    # --- a reference-encoder stream: LZ77_E_STREAM on every rank, nobody hangs --------
    #     (needs >= one block of tokens per rank, else root decodes alone; the 512 KiB
    #     blocks of the large window would need minutes of reference encoding)
    raised = sb > 8191
    if not raised:
        n_ref = 600_000 * world
        ref = None
        if rank == 0:
            ref_bytes = oracle().ref_encode(synth.zipf_text(n_ref, seed=9).numpy(), sb, la)
            ref = torch.frombuffer(bytearray(ref_bytes) + bytearray(16), dtype=torch.uint8).to(dev)
            ref = ref[:len(ref_bytes)]
        try:
            api.decode_sharded_tensor(ref, out_cap=n_ref)
        except api.Lz77Error as e:
            raised = e.rc == api.E_STREAM
    q.put((rank, ok and raised, why or ("" if raised else "reference stream did not fail")))
    dist.barrier()
    api.comm_destroy()
    dist.destroy_process_group()


@pytest.mark.parametrize("sb,la", [(4095, 15), (1000, 20), (65535, 255)])
def test_sharded_codec_nccl(sb, la):
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
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in got), got


_MGPU_SCRIPT = r"""
import sys
sys.path.insert(0, %(root)r)
import numpy as np, torch
import lz77_b200
from lz77_b200 import api, synth
from oracle import oracle
n_gpus = min(torch.cuda.device_count(), 4)
api.mgpu_init(n_gpus)
for sb, la, n in ((4095, 15, 37 * 65536 + 4321), (1000, 20, 21 * 65536 + 77), (65535, 255, 9 * 524288 + 5)):
    data = synth.zipf_text(n, seed=sb).numpy().tobytes()
    merged = api.mgpu_encode(data, la=la, sb=sb)
    assert oracle().decode(merged) == data, (sb, la, "oracle decode")
    assert api.mgpu_decode(merged, n) == data, (sb, la, "mgpu decode")
    lz77_b200.init(0)
    assert lz77_b200.encode(data, la=la, sb=sb) == merged, (sb, la, "single-GPU stream")
api.mgpu_shutdown()
print("mgpu ok", n_gpus)
"""


def test_mgpu_threads_single_process():
    """lz77_mgpu_*: one worker thread per device inside one process."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-c", _MGPU_SCRIPT % {"root": str(ROOT)}],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mgpu ok" in r.stdout, r.stdout + r.stderr
