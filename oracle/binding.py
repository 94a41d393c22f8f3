"""ctypes binding of oracle/liblz77oracle.so and a runner for oracle/_ref/lz77.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "liblz77oracle.so"
_REF = _HERE / "_ref" / "lz77"
REFERENCE_SRC = Path("/root/reference")


def build_oracle(force: bool = False) -> None:
    """Compile the C restatement and (only where /root/reference exists, i.e. in
    the build container) the unmodified reference binary into oracle/_ref/."""
    need_lib = force or not _LIB.exists() or (
        _LIB.stat().st_mtime < (_HERE / "lz77_oracle.c").stat().st_mtime
    )
    need_ref = (REFERENCE_SRC / "lz77.c").exists() and (force or not _REF.exists())
    if need_lib or need_ref:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)


def ref_binary() -> Path | None:
    """Path of the compiled reference CLI, or None when it was never built."""
    return _REF if _REF.exists() else None


def ref_run(mode: str, data: bytes, sb: int | None = None, la: int | None = None,
            timeout: float = 600.0) -> bytes:
    """Run the reference CLI (``-c`` or ``-d``) on ``data`` via tmpfs files."""
    exe = ref_binary()
    if exe is None:
        raise RuntimeError("oracle/_ref/lz77 is not built (run `make -C oracle`)")
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=tmpdir) as d:
        fin, fout = os.path.join(d, "in"), os.path.join(d, "out")
        with open(fin, "wb") as f:
            f.write(data)
        cmd = [str(exe), mode, "-i", fin, "-o", fout]
        if sb is not None:
            cmd += ["-s", str(sb)]
        if la is not None:
            cmd += ["-l", str(la)]
        subprocess.run(cmd, check=True, timeout=timeout, capture_output=True)
        with open(fout, "rb") as f:
            return f.read()


def _u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf, dtype=np.uint8)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


class Oracle:
    def __init__(self) -> None:
        build_oracle()
        lib = C.CDLL(str(_LIB))
        p8 = C.POINTER(C.c_uint8)
        lib.lz77o_bitof.argtypes = [C.c_int]
        lib.lz77o_bitof.restype = C.c_int
        lib.lz77o_token_bits.argtypes = [C.c_int, C.c_int]
        lib.lz77o_token_bits.restype = C.c_int
        lib.lz77o_encode_bound.argtypes = [C.c_long, C.c_int, C.c_int]
        lib.lz77o_encode_bound.restype = C.c_long
        lib.lz77o_ref_encode.argtypes = [p8, C.c_long, C.c_int, C.c_int, p8, C.c_long]
        lib.lz77o_ref_encode.restype = C.c_long
        lib.lz77o_decode.argtypes = [p8, C.c_long, p8, C.c_long]
        lib.lz77o_decode.restype = C.c_long
        lib.lz77o_blocked_encode.argtypes = [p8, C.c_long, C.c_int, C.c_int, C.c_long,
                                             p8, C.c_long, C.POINTER(C.c_long)]
        lib.lz77o_blocked_encode.restype = C.c_long
        lib.lz77o_segmented_encode.argtypes = [p8, C.c_long, C.c_int, C.c_int, C.c_long,
                                               C.c_long, p8, C.c_long, C.POINTER(C.c_long)]
        lib.lz77o_segmented_encode.restype = C.c_long
        lib.lz77o_unpack_tokens.argtypes = [p8, C.c_long, C.POINTER(C.c_int),
                                            C.POINTER(C.c_int), C.POINTER(C.c_int32),
                                            C.POINTER(C.c_int32), p8, C.c_long]
        lib.lz77o_unpack_tokens.restype = C.c_long
        self.lib = lib
        self._p8 = p8

    def _ptr(self, a: np.ndarray):
        return a.ctypes.data_as(self._p8)

    def bitof(self, n: int) -> int:
        return self.lib.lz77o_bitof(n)

    def token_bits(self, sb: int, la: int) -> int:
        return self.lib.lz77o_token_bits(sb, la)

    def encode_bound(self, n: int, sb: int, la: int) -> int:
        return self.lib.lz77o_encode_bound(n, sb, la)

    def ref_encode(self, data, sb: int = -1, la: int = -1) -> bytes:
        src = _u8(data)
        esb = 4095 if sb == -1 else sb
        ela = 15 if la == -1 else la
        cap = self.encode_bound(src.size, max(esb, 1), max(ela, 1)) + 8
        out = np.zeros(cap, dtype=np.uint8)
        n = self.lib.lz77o_ref_encode(self._ptr(src), src.size, sb, la, self._ptr(out), cap)
        if n < 0:
            raise ValueError(f"lz77o_ref_encode failed: {n}")
        return out[:n].tobytes()

    def decode(self, stream) -> bytes:
        src = _u8(stream)
        n = self.lib.lz77o_decode(self._ptr(src), src.size, None, 0)
        if n < 0:
            raise ValueError(f"lz77o_decode failed: {n}")
        out = np.zeros(max(n, 1), dtype=np.uint8)
        m = self.lib.lz77o_decode(self._ptr(src), src.size, self._ptr(out), n)
        if m != n:
            raise ValueError(f"lz77o_decode size mismatch: {m} != {n}")
        return out[:n].tobytes()

    def blocked_encode(self, data, sb: int = -1, la: int = -1, block: int = 0,
                       segment: int = 0):
        """Specification of the GPU encoder (block-independent windows, parse
        restart every ``segment`` bytes); returns (stream, token count)."""
        src = _u8(data)
        esb = 4095 if sb == -1 else sb
        ela = 15 if la == -1 else la
        cap = self.encode_bound(src.size, max(esb, 1), max(ela, 1)) + 8
        out = np.zeros(cap, dtype=np.uint8)
        ntok = C.c_long(0)
        n = self.lib.lz77o_segmented_encode(self._ptr(src), src.size, sb, la, block, segment,
                                            self._ptr(out), cap, C.byref(ntok))
        if n < 0:
            raise ValueError(f"lz77o_segmented_encode failed: {n}")
        return out[:n].tobytes(), ntok.value

    def unpack_tokens(self, stream):
        src = _u8(stream)
        sb, la = C.c_int(0), C.c_int(0)
        k = self.lib.lz77o_unpack_tokens(self._ptr(src), src.size, C.byref(sb), C.byref(la),
                                         None, None, None, 0)
        if k < 0:
            raise ValueError(f"lz77o_unpack_tokens failed: {k}")
        off = np.zeros(max(k, 1), dtype=np.int32)
        ln = np.zeros(max(k, 1), dtype=np.int32)
        nx = np.zeros(max(k, 1), dtype=np.uint8)
        self.lib.lz77o_unpack_tokens(self._ptr(src), src.size, C.byref(sb), C.byref(la),
                                     off.ctypes.data_as(C.POINTER(C.c_int32)),
                                     ln.ctypes.data_as(C.POINTER(C.c_int32)),
                                     self._ptr(nx), k)
        return sb.value, la.value, off[:k], ln[:k], nx[:k]


_ORACLE: Oracle | None = None


def oracle() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE
