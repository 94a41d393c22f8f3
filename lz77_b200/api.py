"""ctypes binding of liblz77b200.so -- the host-side mirror of the reference's
codec interface (``encode(file, out, la, sb)`` / ``decode(file, out)``,
reference lz77.h:14-15) over the C ABI declared in include/lz77_b200.h.

There is no CPU implementation behind these calls: if the CUDA library is not
built, or no CUDA device is visible, they raise.  torch is used only for
device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("LZ77_B200_LIB", _HERE / "liblz77b200.so"))  # override: experiments only

DEFAULT_LA = 15    # reference lz77.c:21
DEFAULT_SB = 4095  # reference lz77.c:22

E_ARG, E_SPACE, E_STREAM, E_NOMEM, E_NODEVICE, E_CUDA, E_COMM = -1, -2, -3, -4, -5, -6, -7
COMM_ID_BYTES = 128

# every symbol include/lz77_b200.h declares
EXPORTS = (
    "lz77_bitof", "lz77_token_bits", "lz77_gpu_encode_bound", "lz77_gpu_block_size",
    "lz77_gpu_segment_size", "lz77_gpu_device_count", "lz77_gpu_init", "lz77_gpu_shutdown",
    "lz77_gpu_strerror", "lz77_gpu_last_error", "lz77_gpu_host_alloc", "lz77_gpu_host_free",
    "lz77_gpu_encode", "lz77_gpu_decode_size", "lz77_gpu_decode", "lz77_gpu_encode_device",
    "lz77_gpu_decode_size_device", "lz77_gpu_decode_device", "lz77_gpu_last_timing",
    "lz77_gpu_set_timing", "lz77_gpu_set_stream", "lz77_gpu_set_host_chunk",
    "lz77_gpu_slice_tokens_device", "lz77_gpu_token_at_device", "lz77_gpu_set_jump_piece",
    "lz77_gpu_set_history", "lz77_gpu_set_fused_pack",
    "lz77_shard_range", "lz77_comm_get_unique_id", "lz77_comm_init", "lz77_comm_destroy",
    "lz77_gpu_encode_sharded_device", "lz77_gpu_decode_sharded_device", "lz77_comm_last_stats",
    "lz77_mgpu_init", "lz77_mgpu_shutdown", "lz77_mgpu_encode", "lz77_mgpu_decode",
)


class Lz77Error(RuntimeError):
    def __init__(self, rc: int, msg: str):
        super().__init__(f"lz77_b200 error {rc}: {msg}")
        self.rc = rc


class Timing(C.Structure):
    _fields_ = [
        ("enc_search_ms", C.c_float), ("enc_scan_ms", C.c_float), ("enc_pack_ms", C.c_float),
        ("dec_scan_ms", C.c_float), ("dec_copy_ms", C.c_float),
        ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
        ("launches", C.c_int), ("n_tokens", C.c_long),
    ]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class CommStats(C.Structure):
    _fields_ = [
        ("sent_bytes", C.c_long), ("recv_bytes", C.c_long), ("scatter_ms", C.c_float),
        ("compute_ms", C.c_float), ("gather_ms", C.c_float), ("total_ms", C.c_float),
        ("collectives", C.c_int),
    ]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def load_library() -> C.CDLL:
    """Load liblz77b200.so (built in-tree by ``__graft_entry__.build()`` or
    ``make -C lz77_b200/csrc``).  Raises if it is missing -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C lz77_b200/csrc` (needs nvcc, sm_100a)")
    lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
    vp, lp, ip = C.c_void_p, C.c_long, C.c_int
    plong = C.POINTER(C.c_long)
    sig = {
        "lz77_bitof": (ip, [ip]),
        "lz77_token_bits": (ip, [ip, ip]),
        "lz77_gpu_encode_bound": (lp, [lp, ip, ip]),
        "lz77_gpu_block_size": (lp, [ip]),
        "lz77_gpu_segment_size": (lp, [ip, ip]),
        "lz77_gpu_device_count": (ip, []),
        "lz77_gpu_init": (ip, [ip]),
        "lz77_gpu_shutdown": (None, []),
        "lz77_gpu_strerror": (C.c_char_p, [ip]),
        "lz77_gpu_last_error": (C.c_char_p, []),
        "lz77_gpu_host_alloc": (vp, [lp]),
        "lz77_gpu_host_free": (None, [vp]),
        "lz77_gpu_encode": (ip, [vp, lp, ip, ip, vp, lp, plong]),
        "lz77_gpu_decode_size": (ip, [vp, lp, plong]),
        "lz77_gpu_decode": (ip, [vp, lp, vp, lp, plong]),
        "lz77_gpu_encode_device": (ip, [vp, lp, ip, ip, vp, lp, plong, plong]),
        "lz77_gpu_decode_size_device": (ip, [vp, lp, plong]),
        "lz77_gpu_decode_device": (ip, [vp, lp, vp, lp, plong]),
        "lz77_gpu_last_timing": (ip, [C.POINTER(Timing)]),
        "lz77_gpu_set_timing": (None, [ip]),
        "lz77_gpu_set_stream": (ip, [vp]),
        "lz77_gpu_set_host_chunk": (None, [lp]),
        "lz77_gpu_slice_tokens_device": (ip, [vp, lp, lp, lp, vp, lp, plong]),
        "lz77_gpu_token_at_device": (ip, [vp, lp, lp, plong, plong]),
        "lz77_gpu_set_jump_piece": (ip, [lp]),
        "lz77_gpu_set_history": (None, [ip]),
        "lz77_gpu_set_fused_pack": (None, [ip]),
        "lz77_shard_range": (ip, [lp, ip, lp, ip, plong, plong]),
        "lz77_comm_get_unique_id": (ip, [vp]),
        "lz77_comm_init": (ip, [vp, ip, ip]),
        "lz77_comm_destroy": (None, []),
        "lz77_gpu_encode_sharded_device": (ip, [vp, lp, ip, ip, vp, lp, plong, plong, ip]),
        "lz77_gpu_decode_sharded_device": (ip, [vp, lp, vp, lp, plong, ip]),
        "lz77_comm_last_stats": (ip, [C.POINTER(CommStats)]),
        "lz77_mgpu_init": (ip, [ip]),
        "lz77_mgpu_shutdown": (None, []),
        "lz77_mgpu_encode": (ip, [vp, lp, ip, ip, vp, lp, plong]),
        "lz77_mgpu_decode": (ip, [vp, lp, vp, lp, plong]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        lib = load_library()
        msg = lib.lz77_gpu_strerror(rc).decode()
        if rc in (E_CUDA, E_COMM, E_STREAM):
            detail = lib.lz77_gpu_last_error().decode()
            if detail:
                msg += ": " + detail
        raise Lz77Error(rc, msg)


# ---- format arithmetic (no device needed) ----------------------------------

def bitof(n: int) -> int:
    return load_library().lz77_bitof(n)


def token_bits(sb: int = DEFAULT_SB, la: int = DEFAULT_LA) -> int:
    return load_library().lz77_token_bits(sb, la)


def encode_bound(n_in: int, sb: int = -1, la: int = -1) -> int:
    return load_library().lz77_gpu_encode_bound(n_in, sb, la)


def block_size(sb: int = -1) -> int:
    return load_library().lz77_gpu_block_size(sb)


def segment_size(sb: int = -1, la: int = -1) -> int:
    """Bytes after which the greedy parse restarts, for these parameters (and the current
    fused-pack setting): part of the encoder's specification."""
    return load_library().lz77_gpu_segment_size(sb, la)


# ---- lifetime ----------------------------------------------------------------

_device = None


def init(device: int | None = None) -> int:
    """Bind the library to one GPU (default: LOCAL_RANK or 0)."""
    global _device
    lib = load_library()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if _device is None else _device
    if lib.lz77_gpu_device_count() <= 0:
        raise Lz77Error(E_NODEVICE, "no CUDA device visible; this codec has no CPU path")
    _check(lib.lz77_gpu_init(device))
    _device = device
    return device


def shutdown() -> None:
    global _device
    if _lib is not None:
        _lib.lz77_gpu_shutdown()
    _device = None


def last_timing() -> dict:
    t = Timing()
    _check(load_library().lz77_gpu_last_timing(C.byref(t)))
    return t.as_dict()


def set_stream(cuda_stream: int | None) -> None:
    """Run later calls on this CUDA stream (e.g. ``torch.cuda.current_stream().cuda_stream``);
    None restores the library's own stream."""
    _check(load_library().lz77_gpu_set_stream(cuda_stream or None))


def set_host_chunk(nbytes: int) -> None:
    """Chunk size of the pipelined host entry points (<= 0: no chunking)."""
    load_library().lz77_gpu_set_host_chunk(nbytes)


def set_timing(enabled: bool) -> None:
    load_library().lz77_gpu_set_timing(1 if enabled else 0)


def set_history(enabled: bool) -> None:
    """Encoder: let the match window slide across block seams like the reference's
    (lz77.c:101-105).  Better ratio at large windows; the stream no longer decodes block
    by block (pointer jumping) and cannot shard for decode."""
    load_library().lz77_gpu_set_history(1 if enabled else 0)


def set_fused_pack(enabled: bool) -> None:
    """Encoder, 24-bit tokens: pack inside the search kernel (no unpacked tokens in HBM)."""
    load_library().lz77_gpu_set_fused_pack(1 if enabled else 0)


def set_jump_piece(nbytes: int) -> None:
    """Output bytes per piece of the pointer-jumping decoder (0: the default, 64 MiB)."""
    _check(load_library().lz77_gpu_set_jump_piece(nbytes))


# ---- host buffers: the call a user of the reference makes ---------------------

def _as_u8(buf):
    import numpy as np
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf.reshape(-1).view(np.uint8))
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def encode(data, la: int = -1, sb: int = -1) -> bytes:
    """``encode(file, out, la, sb)`` of the reference (lz77.c:51) on buffers:
    returns the compressed stream.  ``-1`` selects the defaults."""
    import numpy as np
    init()
    lib = load_library()
    src = _as_u8(data)
    cap = lib.lz77_gpu_encode_bound(src.size, sb, la) + 16
    if cap < 16:
        raise Lz77Error(E_ARG, "bad argument")
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_long(0)
    _check(lib.lz77_gpu_encode(src.ctypes.data, src.size, sb, la, out.ctypes.data, cap,
                               C.byref(n)))
    return out[:n.value].tobytes()


def decode_size(stream) -> int:
    init()
    src = _as_u8(stream)
    n = C.c_long(0)
    _check(load_library().lz77_gpu_decode_size(src.ctypes.data, src.size, C.byref(n)))
    return n.value


def decode(stream) -> bytes:
    """``decode(file, out)`` of the reference (lz77.c:148) on buffers."""
    import numpy as np
    init()
    lib = load_library()
    src = _as_u8(stream)
    n = C.c_long(0)
    _check(lib.lz77_gpu_decode_size(src.ctypes.data, src.size, C.byref(n)))
    out = np.empty(max(n.value, 1), dtype=np.uint8)
    m = C.c_long(0)
    _check(lib.lz77_gpu_decode(src.ctypes.data, src.size, out.ctypes.data, n.value, C.byref(m)))
    return out[:m.value].tobytes()


# ---- pinned host buffers (fast host<->device copies) --------------------------

class PinnedBuffer:
    """Page-locked host memory from the library, exposed as a numpy array."""

    def __init__(self, nbytes: int):
        import numpy as np
        lib = load_library()
        self.nbytes = int(nbytes)
        self.ptr = lib.lz77_gpu_host_alloc(self.nbytes)
        if not self.ptr:
            raise Lz77Error(E_NOMEM, "pinned allocation failed")
        arr_t = C.c_uint8 * max(self.nbytes, 1)
        self.array = np.frombuffer(arr_t.from_address(self.ptr), dtype=np.uint8)[:self.nbytes]

    def free(self) -> None:
        if self.ptr:
            self.array = None
            load_library().lz77_gpu_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def encode_into(src_ptr: int, n_in: int, dst_ptr: int, dst_cap: int, la: int = -1,
                sb: int = -1) -> int:
    """Host-pointer encode (pinned or pageable); returns the stream size."""
    n = C.c_long(0)
    _check(load_library().lz77_gpu_encode(src_ptr, n_in, sb, la, dst_ptr, dst_cap, C.byref(n)))
    return n.value


def decode_into(src_ptr: int, n_in: int, dst_ptr: int, dst_cap: int) -> int:
    n = C.c_long(0)
    _check(load_library().lz77_gpu_decode(src_ptr, n_in, dst_ptr, dst_cap, C.byref(n)))
    return n.value


# ---- device tensors -----------------------------------------------------------

def _round16(n: int) -> int:
    return (n + 15) & ~15


def encode_tensor(src, la: int = -1, sb: int = -1, out=None):
    """Encode a CUDA uint8 tensor; returns (stream tensor view, token count).
    ``out`` may be a preallocated CUDA uint8 tensor of at least
    ``round16(encode_bound(n))`` bytes."""
    import torch
    assert src.is_cuda and src.dtype == torch.uint8 and src.is_contiguous()
    if src.data_ptr() % 16:
        raise Lz77Error(E_ARG, "device buffers must be 16-byte aligned (clone the view)")
    init(src.device.index)
    lib = load_library()
    n_in = src.numel()
    cap = _round16(lib.lz77_gpu_encode_bound(n_in, sb, la))
    if out is None:
        out = torch.empty(cap, dtype=torch.uint8, device=src.device)
    assert out.is_cuda and out.numel() >= cap
    torch.cuda.current_stream(src.device).synchronize()
    n, k = C.c_long(0), C.c_long(0)
    _check(lib.lz77_gpu_encode_device(src.data_ptr(), n_in, sb, la, out.data_ptr(), out.numel(),
                                      C.byref(n), C.byref(k)))
    return out[:n.value], k.value


def decode_size_tensor(stream) -> int:
    import torch
    assert stream.is_cuda and stream.dtype == torch.uint8 and stream.is_contiguous()
    init(stream.device.index)
    torch.cuda.current_stream(stream.device).synchronize()
    n = C.c_long(0)
    _check(load_library().lz77_gpu_decode_size_device(stream.data_ptr(), stream.numel(),
                                                      C.byref(n)))
    return n.value


def slice_tokens_tensor(stream, tok_lo: int, tok_hi: int):
    """Standalone stream (CUDA uint8 tensor) holding tokens [tok_lo, tok_hi) of `stream`."""
    import torch
    assert stream.is_cuda and stream.dtype == torch.uint8 and stream.is_contiguous()
    init(stream.device.index)
    torch.cuda.current_stream(stream.device).synchronize()
    hdr = bytes(stream[:4].cpu().numpy())
    T = token_bits(hdr[0] | hdr[1] << 8, hdr[2] | hdr[3] << 8)
    cap = _round16(4 + ((tok_hi - tok_lo) * T + 7) // 8 + 4)
    out = torch.empty(cap, dtype=torch.uint8, device=stream.device)
    n = C.c_long(0)
    _check(load_library().lz77_gpu_slice_tokens_device(stream.data_ptr(), stream.numel(), tok_lo,
                                                       tok_hi, out.data_ptr(), cap, C.byref(n)))
    return out[:n.value]


def token_at_tensor(stream, pos: int):
    """(index of the token holding decoded byte `pos`, decoded position of its first byte)."""
    import torch
    assert stream.is_cuda and stream.dtype == torch.uint8 and stream.is_contiguous()
    init(stream.device.index)
    torch.cuda.current_stream(stream.device).synchronize()
    k, p = C.c_long(0), C.c_long(0)
    _check(load_library().lz77_gpu_token_at_device(stream.data_ptr(), stream.numel(), pos,
                                                   C.byref(k), C.byref(p)))
    return k.value, p.value


def decode_tensor(stream, out=None):
    """Decode a CUDA uint8 stream tensor into a CUDA uint8 tensor."""
    import torch
    assert stream.is_cuda and stream.dtype == torch.uint8 and stream.is_contiguous()
    init(stream.device.index)
    lib = load_library()
    torch.cuda.current_stream(stream.device).synchronize()
    if out is None:
        n = decode_size_tensor(stream)
        out = torch.empty(_round16(max(n, 1)), dtype=torch.uint8, device=stream.device)
    m = C.c_long(0)
    _check(lib.lz77_gpu_decode_device(stream.data_ptr(), stream.numel(), out.data_ptr(),
                                      out.numel(), C.byref(m)))
    return out[:m.value]


# ---- several GPUs: one input, one stream (NCCL inside the library) -------------

def shard_range(n_bytes: int, world: int, block: int, rank: int):
    """Byte range [lo, hi) of `rank`: a contiguous run of whole blocks (host arithmetic)."""
    lo, hi = C.c_long(0), C.c_long(0)
    _check(load_library().lz77_shard_range(n_bytes, world, block, rank, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def comm_unique_id() -> bytes:
    """The id rank 0 hands to every rank (any transport) before ``comm_init``."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(load_library().lz77_comm_get_unique_id(buf))
    return buf.raw


def comm_init(uid: bytes, rank: int, world: int) -> None:
    """Join the library's NCCL communicator (after ``init(device)``)."""
    assert len(uid) == COMM_ID_BYTES
    buf = C.create_string_buffer(uid, COMM_ID_BYTES)
    _check(load_library().lz77_comm_init(buf, rank, world))


def comm_init_torch(device) -> None:
    """``comm_init`` with the id carried by torch.distributed's default group."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    init(device.index)
    t = torch.zeros(COMM_ID_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    comm_init(bytes(t.cpu().numpy()), rank, world)


def comm_destroy() -> None:
    if _lib is not None:
        _lib.lz77_comm_destroy()


def comm_stats() -> dict:
    s = CommStats()
    _check(load_library().lz77_comm_last_stats(C.byref(s)))
    return s.as_dict()


def encode_sharded_tensor(src, n_in: int = 0, la: int = -1, sb: int = -1, out=None, root: int = 0):
    """Collective: every rank calls it; on root `src` is the whole input (CUDA uint8) and
    the result is (merged stream view, total tokens); the other ranks pass ``None`` and
    get (None, total tokens)."""
    import torch
    lib = load_library()
    n, k = C.c_long(0), C.c_long(0)
    if src is not None:
        assert src.is_cuda and src.dtype == torch.uint8 and src.is_contiguous()
        n_in = src.numel()
        cap = _round16(lib.lz77_gpu_encode_bound(n_in, sb, la))
        if out is None:
            out = torch.empty(cap, dtype=torch.uint8, device=src.device)
        torch.cuda.current_stream(src.device).synchronize()
        _check(lib.lz77_gpu_encode_sharded_device(src.data_ptr(), n_in, sb, la, out.data_ptr(),
                                                  out.numel(), C.byref(n), C.byref(k), root))
        return out[:n.value], k.value
    _check(lib.lz77_gpu_encode_sharded_device(None, 0, sb, la, None, 0, C.byref(n), C.byref(k),
                                              root))
    return None, k.value


def decode_sharded_tensor(stream, out=None, out_cap: int = 0, root: int = 0):
    """Collective: root passes the stream (CUDA uint8) and an output tensor (or `out_cap`
    bytes to allocate); returns the decoded view on root, None elsewhere."""
    import torch
    lib = load_library()
    m = C.c_long(0)
    if stream is not None:
        assert stream.is_cuda and stream.dtype == torch.uint8 and stream.is_contiguous()
        if out is None:
            out = torch.empty(_round16(max(out_cap, 16)), dtype=torch.uint8, device=stream.device)
        torch.cuda.current_stream(stream.device).synchronize()
        _check(lib.lz77_gpu_decode_sharded_device(stream.data_ptr(), stream.numel(), out.data_ptr(),
                                                  out.numel(), C.byref(m), root))
        return out[:m.value]
    _check(lib.lz77_gpu_decode_sharded_device(None, 0, None, 0, C.byref(m), root))
    return None


def mgpu_init(n_gpus: int) -> None:
    _check(load_library().lz77_mgpu_init(n_gpus))


def mgpu_shutdown() -> None:
    if _lib is not None:
        _lib.lz77_mgpu_shutdown()


def mgpu_encode(data, la: int = -1, sb: int = -1) -> bytes:
    """One process, every GPU of ``mgpu_init``: host buffer in, merged stream out."""
    import numpy as np
    lib = load_library()
    src = _as_u8(data)
    cap = lib.lz77_gpu_encode_bound(src.size, sb, la) + 16
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_long(0)
    _check(lib.lz77_mgpu_encode(src.ctypes.data, src.size, sb, la, out.ctypes.data, cap, C.byref(n)))
    return out[:n.value].tobytes()


def mgpu_decode(stream, n_out_max: int) -> bytes:
    import numpy as np
    lib = load_library()
    src = _as_u8(stream)
    out = np.empty(max(n_out_max, 1) + 16, dtype=np.uint8)
    m = C.c_long(0)
    _check(lib.lz77_mgpu_decode(src.ctypes.data, src.size, out.ctypes.data, out.size, C.byref(m)))
    return out[:m.value].tobytes()
