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

E_ARG, E_SPACE, E_STREAM, E_NOMEM, E_NODEVICE, E_CUDA = -1, -2, -3, -4, -5, -6

# every symbol include/lz77_b200.h declares
EXPORTS = (
    "lz77_bitof", "lz77_token_bits", "lz77_gpu_encode_bound", "lz77_gpu_block_size",
    "lz77_gpu_segment_size", "lz77_gpu_device_count", "lz77_gpu_init", "lz77_gpu_shutdown",
    "lz77_gpu_strerror", "lz77_gpu_last_error", "lz77_gpu_host_alloc", "lz77_gpu_host_free",
    "lz77_gpu_encode", "lz77_gpu_decode_size", "lz77_gpu_decode", "lz77_gpu_encode_device",
    "lz77_gpu_decode_size_device", "lz77_gpu_decode_device", "lz77_gpu_last_timing",
    "lz77_gpu_set_timing", "lz77_gpu_set_stream", "lz77_gpu_set_host_chunk",
    "lz77_gpu_slice_tokens_device", "lz77_gpu_token_at_device",
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
        "lz77_gpu_segment_size": (lp, []),
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
        if rc == E_CUDA:
            msg += ": " + lib.lz77_gpu_last_error().decode()
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


def segment_size() -> int:
    return load_library().lz77_gpu_segment_size()


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
