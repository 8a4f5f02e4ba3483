"""ctypes front-ends for the two CPU checkers used by the tests (never by the product path).

`Oracle`  -> oracle/liboracle.so        (this repo's C restatement, always available after build())
`Ref`     -> oracle/_ref/libnblic_ref.so (the unmodified reference, when it was compiled here)

Both expose the same four calls over numpy arrays / bytes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libnblic_ref.so")

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_ip = C.POINTER(C.c_int)


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)


def _out_capacity(h: int, w: int) -> int:
    return 2 * h * w + (1 << 16)


class Oracle:
    name = "oracle"

    def __init__(self) -> None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.oracle_q_encode.argtypes = [_u8p, C.c_int, C.c_int, _u16p]
        self.lib.oracle_q_decode.argtypes = [_u16p, C.c_long, _u8p, _ip, _ip]
        self.lib.oracle_n_encode.argtypes = [_u8p, C.c_int, C.c_int, _ip, _ip, _u8p]
        self.lib.oracle_n_decode.argtypes = [_u8p, C.c_long, _u8p, _ip, _ip, _ip, _ip]

    def q_encode(self, img: np.ndarray) -> bytes:
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        out = np.zeros(_out_capacity(h, w) // 2, dtype=np.uint16)
        n = self.lib.oracle_q_encode(img.ctypes.data_as(_u8p), h, w, out.ctypes.data_as(_u16p))
        if n < 0:
            raise ValueError("q_encode failed")
        return out[:n].tobytes()

    def q_decode(self, data: bytes, pixels_hint: int = 100_000_000):
        buf = np.frombuffer(data + b"\0" * (len(data) & 1), dtype=np.uint16).copy()
        hh, ww = C.c_int(-1), C.c_int(-1)
        # header first so the pixel buffer can be sized
        if len(buf) < 4 or buf[0] != 0x3051 or buf[1] != 0x322E:
            return None
        h, w = int(buf[2]), int(buf[3])
        img = np.zeros(max(h * w, 1), dtype=np.uint8)
        rc = self.lib.oracle_q_decode(buf.ctypes.data_as(_u16p), len(buf), img.ctypes.data_as(_u8p), C.byref(hh), C.byref(ww))
        if rc != 0:
            return None
        return img[: h * w].reshape(h, w)

    def n_encode(self, img: np.ndarray, near: int, effort: int):
        """-> (bytes, reconstruction, near_used, effort_used)"""
        work = np.ascontiguousarray(img, dtype=np.uint8).copy()
        h, w = work.shape
        out = np.zeros(_out_capacity(h, w), dtype=np.uint8)
        n_, e_ = C.c_int(near), C.c_int(effort)
        n = self.lib.oracle_n_encode(work.ctypes.data_as(_u8p), h, w, C.byref(n_), C.byref(e_), out.ctypes.data_as(_u8p))
        if n < 0:
            raise ValueError("n_encode failed")
        return out[:n].tobytes(), work, n_.value, e_.value

    def n_decode(self, data: bytes):
        """-> (img, near, effort) or None"""
        if len(data) < 16 or data[:8] != b"NBLIC0.3":
            return None
        h, w = (data[9] << 8) | data[10], (data[11] << 8) | data[12]
        buf = np.frombuffer(data, dtype=np.uint8).copy()
        img = np.zeros(max(h * w, 1), dtype=np.uint8)
        hh, ww, nn, ee = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rc = self.lib.oracle_n_decode(buf.ctypes.data_as(_u8p), len(buf), img.ctypes.data_as(_u8p),
                                      C.byref(hh), C.byref(ww), C.byref(nn), C.byref(ee))
        if rc != 0:
            return None
        return img[: h * w].reshape(h, w), nn.value, ee.value


class Ref:
    """The unmodified reference (NBLIC.h:54,72 / QNBLIC.h:14-18) through ctypes."""
    name = "reference"

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self) -> None:
        self.lib = C.CDLL(REF_SO)
        self.lib.QNBLICcompress.argtypes = [_u16p, _u8p, C.c_int, C.c_int]
        self.lib.QNBLICdecompress.argtypes = [_u16p, _u8p, _ip, _ip]
        self.lib.NBLICcompress.argtypes = [C.c_int, _u8p, _u8p, C.c_int, C.c_int, _ip, _ip]
        self.lib.NBLICdecompress.argtypes = [C.c_int, _u8p, _u8p, _ip, _ip, _ip, _ip]

    def q_encode(self, img: np.ndarray) -> bytes:
        img = np.ascontiguousarray(img, dtype=np.uint8).copy()
        h, w = img.shape
        out = np.zeros(_out_capacity(h, w) // 2, dtype=np.uint16)
        n = self.lib.QNBLICcompress(out.ctypes.data_as(_u16p), img.ctypes.data_as(_u8p), h, w)
        if n < 0:
            raise ValueError("QNBLICcompress failed")
        return out[:n].tobytes()

    def q_decode(self, data: bytes):
        pad = np.zeros(len(data) // 2 + 4096, dtype=np.uint16)
        pad[: len(data) // 2] = np.frombuffer(data[: len(data) & ~1], dtype=np.uint16)
        if len(pad) < 4 or pad[0] != 0x3051 or pad[1] != 0x322E:
            return None
        h, w = int(pad[2]), int(pad[3])
        img = np.zeros(max(h * w, 1), dtype=np.uint8)
        hh, ww = C.c_int(), C.c_int()
        rc = self.lib.QNBLICdecompress(pad.ctypes.data_as(_u16p), img.ctypes.data_as(_u8p), C.byref(hh), C.byref(ww))
        if rc != 0:
            return None
        return img[: h * w].reshape(h, w)

    def n_encode(self, img: np.ndarray, near: int, effort: int):
        work = np.ascontiguousarray(img, dtype=np.uint8).copy()
        h, w = work.shape
        out = np.zeros(_out_capacity(h, w), dtype=np.uint8)
        n_, e_ = C.c_int(near), C.c_int(effort)
        n = self.lib.NBLICcompress(0, out.ctypes.data_as(_u8p), work.ctypes.data_as(_u8p), h, w, C.byref(n_), C.byref(e_))
        if n < 0:
            raise ValueError("NBLICcompress failed")
        return out[:n].tobytes(), work, n_.value, e_.value

    def n_decode(self, data: bytes):
        if len(data) < 16 or data[:8] != b"NBLIC0.3":
            return None
        h, w = (data[9] << 8) | data[10], (data[11] << 8) | data[12]
        buf = np.zeros(len(data) + 4096, dtype=np.uint8)
        buf[: len(data)] = np.frombuffer(data, dtype=np.uint8)
        img = np.zeros(max(h * w, 1), dtype=np.uint8)
        hh, ww, nn, ee = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rc = self.lib.NBLICdecompress(0, buf.ctypes.data_as(_u8p), img.ctypes.data_as(_u8p),
                                      C.byref(hh), C.byref(ww), C.byref(nn), C.byref(ee))
        if rc != 0:
            return None
        return img[: h * w].reshape(h, w), nn.value, ee.value


def load_bmp_gray(path: str) -> np.ndarray:
    """8-bit palettised gray BMP, bottom-up, rows padded to 4 bytes (layout per FileIO.c:170-245)."""
    import struct
    d = open(path, "rb").read()
    off = struct.unpack_from("<I", d, 10)[0]
    w, h = struct.unpack_from("<ii", d, 18)
    stride = (w + 3) & ~3
    rows = np.frombuffer(d, dtype=np.uint8, count=stride * abs(h), offset=off).reshape(abs(h), stride)[:, :w]
    return np.ascontiguousarray(rows[::-1] if h > 0 else rows)
