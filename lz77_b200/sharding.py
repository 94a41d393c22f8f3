"""Block-range sharding of one input across ranks (one process per GPU).

The encoder cuts its input into independent blocks (``block_size(sb)`` bytes: no
match crosses a block boundary and a token starts on every boundary), so a rank
can encode any run of whole blocks on its own and the token payloads of
consecutive shards concatenate into exactly the stream a single GPU would have
written for the whole input.  Nothing inside the codec communicates; the only
exchange is moving buffers:

  scatter_input    root -> ranks : each rank's run of whole blocks
  gather_stream    ranks -> root : variable-length token payloads (+ a tiny
                                   all-gather of token counts for the offsets)

Both work on whatever device the process group's backend moves (CUDA tensors
over NCCL / NVLink on the GPU box, CPU tensors over gloo in the CPU tests).
The codec itself is passed in as a callable, so this module holds no compute.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch

HEADER_BYTES = 4  # SB:16 LA:16, reference lz77.c:74-75


def shard_ranges(n_bytes: int, world: int, block: int) -> List[Tuple[int, int]]:
    """Byte range [lo, hi) of every rank: contiguous runs of whole blocks, as even
    as whole blocks allow (earlier ranks take the extra block); the last
    non-empty range ends at n_bytes."""
    if world < 1 or block < 1 or n_bytes < 0:
        raise ValueError("bad shard arguments")
    n_blocks = (n_bytes + block - 1) // block
    base, extra = divmod(n_blocks, world)
    out, lo_blk = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        lo = min(lo_blk * block, n_bytes)
        hi = min((lo_blk + cnt) * block, n_bytes)
        out.append((lo, hi))
        lo_blk += cnt
    return out


def payload_bits(n_tokens: int, token_bits: int) -> int:
    return n_tokens * token_bits


def split_stream(stream: np.ndarray) -> Tuple[bytes, np.ndarray]:
    """(4-byte header, token payload bytes) of a stream."""
    s = np.ascontiguousarray(stream, dtype=np.uint8)
    return s[:HEADER_BYTES].tobytes(), s[HEADER_BYTES:]


def merge_payloads(header: bytes, payloads: Sequence[np.ndarray], n_tokens: Sequence[int],
                   token_bits: int) -> np.ndarray:
    """Concatenate the token payloads of consecutive shards into one stream.

    Token k of the merged stream sits at bit 32 + k*T (reference lz77.c:246-252 /
    bitio.c:203-239, LSB first).  When T is a multiple of 8 (both benchmark
    parameter sets: 24 and 32 bits) the payloads are byte strings and simply
    concatenate; otherwise every shard after the first starts inside a byte and
    is shifted into place bit-exactly (the <= 7-bit seam merge)."""
    assert len(payloads) == len(n_tokens)
    total_tokens = int(sum(n_tokens))
    total_bits = total_tokens * token_bits
    out = np.zeros(HEADER_BYTES + (total_bits + 7) // 8, dtype=np.uint8)
    out[:HEADER_BYTES] = np.frombuffer(header, dtype=np.uint8)
    if token_bits % 8 == 0:
        pos = HEADER_BYTES
        for p, k in zip(payloads, n_tokens):
            nb = k * token_bits // 8
            out[pos:pos + nb] = np.asarray(p[:nb], dtype=np.uint8)
            pos += nb
        return out
    bit_pos = 0
    body = out[HEADER_BYTES:]
    for p, k in zip(payloads, n_tokens):
        nbits = k * token_bits
        if nbits == 0:
            continue
        src = np.asarray(p[:(nbits + 7) // 8], dtype=np.uint8)
        bits = np.unpackbits(src, bitorder="little")[:nbits]
        shift = bit_pos & 7
        padded = np.concatenate([np.zeros(shift, dtype=np.uint8), bits])
        packed = np.packbits(padded, bitorder="little")
        first = bit_pos >> 3
        body[first:first + packed.size] |= packed  # the seam byte is OR-merged
        bit_pos += nbits
    return out


# --------------------------------------------------------------------------- #
# collectives (torch.distributed; NCCL on the GPU box, gloo in the CPU tests)
# --------------------------------------------------------------------------- #

def _dist():
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    return dist


def scatter_input(data: torch.Tensor | None, n_bytes: int, block: int, device,
                  root: int = 0) -> torch.Tensor:
    """Root holds `data` (uint8, n_bytes); every rank returns its run of whole
    blocks.  Grouped point-to-point sends (the shards differ in size)."""
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    ranges = shard_ranges(n_bytes, world, block)
    lo, hi = ranges[rank]
    mine = torch.empty(hi - lo, dtype=torch.uint8, device=device)
    if rank == root:
        reqs = []
        for r, (a, b) in enumerate(ranges):
            if r == root:
                mine.copy_(data[a:b])
            elif b > a:
                reqs.append(dist.isend(data[a:b].contiguous(), dst=r))
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(mine, src=root)
    return mine


def gather_stream(local_stream: torch.Tensor, n_tokens: int, token_bits: int, root: int = 0):
    """Every rank passes the stream its shard encoded to (header + payload, uint8
    tensor on the backend's device); root returns the merged stream as a numpy
    array, the other ranks None.  One all-gather of the token counts gives every
    shard its bit offset 32 + T * sum(K_before); the payloads then travel with
    point-to-point sends."""
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = local_stream.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([n_tokens], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine)
    counts = [int(c) for c in counts.cpu().tolist()]
    nbytes = [(c * token_bits + 7) // 8 for c in counts]
    if rank != root:
        if nbytes[rank] > 0:
            dist.send(local_stream[HEADER_BYTES:HEADER_BYTES + nbytes[rank]].contiguous(),
                      dst=root)
        return None
    payloads = []
    for r in range(world):
        if r == root:
            payloads.append(local_stream[HEADER_BYTES:HEADER_BYTES + nbytes[r]].cpu().numpy())
        else:
            buf = torch.empty(nbytes[r], dtype=torch.uint8, device=dev)
            if nbytes[r] > 0:
                dist.recv(buf, src=r)
            payloads.append(buf.cpu().numpy())
    header = local_stream[:HEADER_BYTES].cpu().numpy().tobytes()
    return merge_payloads(header, payloads, counts, token_bits)


def encode_sharded(data: torch.Tensor | None, n_bytes: int, sb: int, la: int, block: int,
                   token_bits: int, encode_fn: Callable, device, root: int = 0):
    """scatter -> per-rank encode -> gather.  `encode_fn(shard_tensor, sb, la)`
    returns (stream tensor, token count) -- on the GPU box that is
    ``lz77_b200.encode_tensor``.  Root returns the merged stream (numpy)."""
    shard = scatter_input(data, n_bytes, block, device, root)
    stream, k = encode_fn(shard, sb, la)
    return gather_stream(stream, k, token_bits, root)


# --------------------------------------------------------------------------- #
# sharded decode of ONE block-structured stream
# --------------------------------------------------------------------------- #

class DecodeCodec:
    """The four operations decode_sharded needs, as callables on uint8 tensors of the
    backend's device (on the GPU box: the ``lz77_b200`` functions of the same names):

      slice_tokens(stream, a, b) -> standalone stream with tokens [a, b)
      decode_size(stream)        -> decoded size
      token_at(stream, pos)      -> (token holding decoded byte pos, its decoded position)
      decode(stream)             -> decoded bytes
    """

    def __init__(self, slice_tokens, decode_size, token_at, decode):
        self.slice_tokens, self.decode_size = slice_tokens, decode_size
        self.token_at, self.decode = token_at, decode


def token_slices(n_tokens: int, world: int) -> List[Tuple[int, int]]:
    """Even split of the token array: [lo, hi) per rank."""
    return [(n_tokens * r // world, n_tokens * (r + 1) // world) for r in range(world)]


def decode_sharded(stream: torch.Tensor | None, block: int, token_bits: int, codec: DecodeCodec,
                   device, root: int = 0):
    """Decode one stream of the block-parallel encoder on all ranks.

    Tokens are fixed width, so the token array splits evenly without parsing:
      1. root sends rank r the tokens [K r/G, K (r+1)/G) plus `block` tokens of margin
         (a standalone stream each; a token decodes to >= 1 byte, so the margin
         reaches the next block boundary);
      2. every rank sums len+1 over its own tokens; one all-gather of the sums gives
         every slice its decoded position;
      3. every rank looks up the token that starts the first block at or after its
         position (a token starts on every block boundary); one all-gather of these
         split points nudges the slice boundaries to block boundaries;
      4. every rank decodes its run of whole blocks -- independent by construction --
         and the outputs travel to root with point-to-point sends.
    Root returns the decoded bytes (uint8 tensor on `device`), the other ranks None.
    Streams of the reference encoder are not block-structured (a match may reach
    back SB bytes from anywhere): replicas only, ValueError here -- on EVERY rank, after
    the collective in which the ranks compare notes (a rank that raised on its own would
    leave the others waiting in the next all-gather).  The same holds for a codec error
    on one rank: it travels as a flag and every rank raises."""
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == root:
        n_tokens = ((stream.numel() - HEADER_BYTES) * 8) // token_bits
        meta[0], meta[1] = n_tokens, stream.numel()
    dist.broadcast(meta, src=root)
    n_tokens = int(meta[0])
    if n_tokens // world < block:
        # fewer than one block of tokens per rank: not worth a split
        if rank != root:
            return None
        return codec.decode(stream)

    slices = token_slices(n_tokens, world)
    with_margin = [(lo, min(n_tokens, hi + block)) for lo, hi in slices]
    nbytes = [HEADER_BYTES + ((b - a) * token_bits + 7) // 8 for a, b in with_margin]
    # 1. token slices (with margin) from root
    if rank == root:
        reqs, local = [], None
        for r, (a, b) in enumerate(with_margin):
            sub = codec.slice_tokens(stream, a, b)
            if r == root:
                local = sub
            else:
                reqs.append(dist.isend(sub.contiguous(), dst=r))
        for q in reqs:
            q.wait()
    else:
        local = torch.empty(nbytes[rank], dtype=torch.uint8, device=device)
        dist.recv(local, src=root)
    k_lo, k_hi = slices[rank]

    def gather_i64(value: int) -> List[int]:
        got = torch.zeros(world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(got, torch.tensor([value], dtype=torch.int64, device=device))
        return [int(v) for v in got.cpu().tolist()]

    # 2. decoded size of the own tokens -> decoded position of every slice
    #    (-1 = this rank's codec failed; every rank sees it in the same all-gather)
    try:
        own = codec.decode_size(codec.slice_tokens(local, 0, k_hi - k_lo))
    except Exception:  # noqa: BLE001 -- reported collectively below
        own = -1
    sums = gather_i64(own)
    if min(sums) < 0:
        raise ValueError(f"rank {sums.index(min(sums))} could not read its token slice")
    pos = sum(sums[:rank])
    # 3. first block boundary at or after the slice start -> split token (-1 = no token
    #    starts there: not a stream of the block encoder)
    to_boundary = (-pos) % block
    try:
        k_rel, k_pos = codec.token_at(local, to_boundary)
        split = k_lo + k_rel if k_pos == to_boundary else -1
    except Exception:  # noqa: BLE001
        split = -1
    splits = gather_i64(split)
    if min(splits) < 0:
        raise ValueError("no token starts on the block boundary: not a stream of the block encoder")
    splits = splits + [n_tokens]
    a, b = splits[rank] - k_lo, splits[rank + 1] - k_lo
    # 4. decode the run of whole blocks, gather at root
    try:
        part = codec.decode(codec.slice_tokens(local, a, b)) if b > a else \
            torch.empty(0, dtype=torch.uint8, device=device)
        size = part.numel()
    except Exception:  # noqa: BLE001
        part, size = None, -1
    sizes = gather_i64(size)
    if min(sizes) < 0:
        raise ValueError(f"rank {sizes.index(min(sizes))} could not decode its blocks")
    if rank != root:
        if sizes[rank] > 0:
            dist.send(part.contiguous(), dst=root)
        return None
    out = torch.empty(sum(sizes), dtype=torch.uint8, device=device)
    at = 0
    for r in range(world):
        if r == root:
            out[at:at + sizes[r]] = part
        elif sizes[r] > 0:
            dist.recv(out[at:at + sizes[r]], src=r)
        at += sizes[r]
    return out
