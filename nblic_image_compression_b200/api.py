"""ctypes mirror of include/nblic_b200.h (no torch types cross this boundary: pointers and sizes only).

`Codec` is the batch interface; `legacy` exposes the reference's own five entry points
(src/NBLIC.h:54,72, src/QNBLIC.h:14-18) exactly as the reference CLI would call them.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .build import LIB, LIB_SEQ

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_ip = C.POINTER(C.c_int)
_szp = C.POINTER(C.c_size_t)
_u64p = C.POINTER(C.c_uint64)
_pp = C.POINTER(C.c_void_p)

MAP_AUTO, MAP_WARP, MAP_LANE, MAP_WARP4 = 0, 1, 2, 3
PIPE_AUTO, PIPE_NEVER, PIPE_ALWAYS = -1, 0, 1
OK, BAD_DIMS, BAD_HEADER, OVERFLOW, CORRUPT = 0, 1, 2, 3, 4

SYMBOLS = [
    # drop-in layer
    "NBLICcompress", "NBLICdecompress", "QNBLICcompress", "QNBLICcompressMultiThread", "QNBLICdecompress",
    "nblic_b200_hint_input_len", "nblic_b200_stream_bound",
    # batch layer
    "nblic_b200_create", "nblic_b200_destroy", "nblic_b200_last_error", "nblic_b200_set_mapping", "nblic_b200_set_pipeline",
    "nblic_b200_encode_batch", "nblic_b200_decode_batch", "nblic_b200_peek", "nblic_b200_host_alloc", "nblic_b200_host_free",
    "nblic_b200_encode_batch_device", "nblic_b200_decode_batch_device", "nblic_b200_synth_gray", "nblic_b200_synth_gray_batch", "nblic_b200_debug_divcheck",
    "nblic_b200_launch_count", "nblic_b200_last_coder_ms", "nblic_b200_last_slots", "nblic_b200_last_mapping", "nblic_b200_stream_handle", "nblic_b200_version",
]

_libs = {}


def load_library(sequential: bool = False) -> C.CDLL:
    """Loads libnblic_b200.so (sequential=True: the test build libnblic_b200_seq.so, which adds the one-agent-per-stream
    kernels behind MAP_LANE); raises if it has not been built (no fallback of any kind)."""
    if sequential in _libs:
        return _libs[sequential]
    path = LIB_SEQ if sequential else LIB
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m nblic_image_compression_b200.build` "
                           "(or __graft_entry__.build()); this package has no CPU fallback")
    lib = C.CDLL(path)
    lib.nblic_b200_create.restype = C.c_void_p
    lib.nblic_b200_create.argtypes = [C.c_int]
    lib.nblic_b200_destroy.argtypes = [C.c_void_p]
    lib.nblic_b200_last_error.restype = C.c_char_p
    lib.nblic_b200_last_error.argtypes = [C.c_void_p]
    lib.nblic_b200_set_mapping.argtypes = [C.c_void_p, C.c_int]
    lib.nblic_b200_set_pipeline.argtypes = [C.c_void_p, C.c_int]
    lib.nblic_b200_stream_bound.restype = C.c_size_t
    lib.nblic_b200_stream_bound.argtypes = [C.c_int, C.c_int]
    lib.nblic_b200_hint_input_len.argtypes = [C.c_size_t]
    lib.nblic_b200_encode_batch.argtypes = [C.c_void_p, C.c_int, _pp, _ip, _ip, C.c_int, C.c_int, _pp, _szp, _szp, _pp, _ip]
    lib.nblic_b200_decode_batch.argtypes = [C.c_void_p, C.c_int, _pp, _szp, _pp, _szp, _ip, _ip, _ip, _ip, _ip]
    lib.nblic_b200_peek.argtypes = [C.c_void_p, C.c_size_t, _ip, _ip, _ip, _ip]
    lib.nblic_b200_encode_batch_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, _u64p, _ip, _ip, C.c_int, C.c_int,
                                                   C.c_void_p, C.c_uint64, _u64p, C.c_void_p, _ip]
    lib.nblic_b200_decode_batch_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, _u64p, C.c_void_p, _u64p, _u64p, _ip]
    lib.nblic_b200_synth_gray.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_int32)]
    lib.nblic_b200_synth_gray_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32)]
    lib.nblic_b200_last_slots.argtypes = [C.c_void_p]
    lib.nblic_b200_debug_divcheck.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.nblic_b200_launch_count.restype = C.c_uint64
    lib.nblic_b200_launch_count.argtypes = [C.c_void_p]
    lib.nblic_b200_last_coder_ms.restype = C.c_float
    lib.nblic_b200_last_coder_ms.argtypes = [C.c_void_p]
    lib.nblic_b200_last_mapping.restype = C.c_char_p
    lib.nblic_b200_last_mapping.argtypes = [C.c_void_p]
    lib.nblic_b200_stream_handle.restype = C.c_void_p
    lib.nblic_b200_stream_handle.argtypes = [C.c_void_p]
    lib.nblic_b200_version.restype = C.c_char_p
    lib.NBLICcompress.argtypes = [C.c_int, _u8p, _u8p, C.c_int, C.c_int, _ip, _ip]
    lib.NBLICdecompress.argtypes = [C.c_int, _u8p, _u8p, _ip, _ip, _ip, _ip]
    lib.QNBLICcompress.argtypes = [_u16p, _u8p, C.c_int, C.c_int]
    lib.QNBLICcompressMultiThread.argtypes = [_u16p, _u8p, C.c_int, C.c_int]
    lib.QNBLICdecompress.argtypes = [_u16p, _u8p, _ip, _ip]
    _libs[sequential] = lib
    return lib


def stream_bound(h: int, w: int) -> int:
    return int(load_library().nblic_b200_stream_bound(h, w))


def peek(data: bytes) -> Optional[Tuple[int, int, int, int]]:
    """(height, width, near, effort) of a stream header, effort 0 = "Q0.2"; None if not a valid header."""
    lib = load_library()
    h, w, n, e = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    buf = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data if data else b"\0")
    rc = lib.nblic_b200_peek(C.cast(buf, C.c_void_p), len(data), C.byref(h), C.byref(w), C.byref(n), C.byref(e))
    return (h.value, w.value, n.value, e.value) if rc == 0 else None


def _ptr_array(arrs: Sequence[Optional[np.ndarray]]):
    out = (C.c_void_p * max(len(arrs), 1))()
    for i, a in enumerate(arrs):
        out[i] = None if a is None else a.ctypes.data
    return out


class Codec:
    """One context on one GPU (nblic_b200_create)."""

    def __init__(self, device: int = 0, mapping: int = MAP_AUTO, sequential: bool = False):
        self.lib = load_library(sequential)
        self.ctx = self.lib.nblic_b200_create(device)
        if not self.ctx:
            raise RuntimeError("nblic_b200_create failed: " + self.lib.nblic_b200_last_error(None).decode())
        self.device = device
        if mapping != MAP_AUTO:
            self.set_mapping(mapping)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.nblic_b200_destroy(self.ctx)
            self.ctx = None

    __del__ = close

    def _err(self) -> str:
        return self.lib.nblic_b200_last_error(self.ctx).decode()

    def set_mapping(self, mapping: int):
        if self.lib.nblic_b200_set_mapping(self.ctx, mapping) != 0:
            raise ValueError("bad mapping")

    def set_pipeline(self, mode: int):
        """PIPE_AUTO (-1) / PIPE_NEVER (0) / PIPE_ALWAYS (1): whole-GPU single-image encode of lossless effort 0 / 1."""
        if self.lib.nblic_b200_set_pipeline(self.ctx, mode) != 0:
            raise ValueError("bad pipeline mode")

    @property
    def launches(self) -> int:
        return int(self.lib.nblic_b200_launch_count(self.ctx))

    @property
    def last_coder_ms(self) -> float:
        return float(self.lib.nblic_b200_last_coder_ms(self.ctx))

    @property
    def stream_handle(self) -> int:
        return int(self.lib.nblic_b200_stream_handle(self.ctx) or 0)

    @property
    def last_mapping(self) -> str:
        return self.lib.nblic_b200_last_mapping(self.ctx).decode()

    # ---- host-buffer batch calls -------------------------------------------------------------
    def encode_batch(self, images: Sequence[np.ndarray], near: int = 0, effort: int = 1, want_recon: bool = False,
                     outs: Optional[List[np.ndarray]] = None):
        """-> (streams: list[bytes | None], recon: list[np.ndarray | None], status: list[int])"""
        n = len(images)
        imgs = [np.ascontiguousarray(a, dtype=np.uint8) for a in images]
        hs = (C.c_int * max(n, 1))(*[a.shape[0] if a.ndim == 2 else 0 for a in imgs])
        ws = (C.c_int * max(n, 1))(*[a.shape[1] if a.ndim == 2 else 0 for a in imgs])
        if outs is None:
            outs = [np.empty(stream_bound(a.shape[0], a.shape[1]), dtype=np.uint8) for a in imgs]
        caps = (C.c_size_t * max(n, 1))(*[o.size for o in outs])
        lens = (C.c_size_t * max(n, 1))()
        status = (C.c_int * max(n, 1))()
        recs = [np.empty_like(a) if want_recon else None for a in imgs]
        rc = self.lib.nblic_b200_encode_batch(self.ctx, n, _ptr_array(imgs), hs, ws, near, effort, _ptr_array(outs), caps, lens,
                                              _ptr_array(recs) if want_recon else None, status)
        if rc < 0:
            raise RuntimeError("nblic_b200_encode_batch: " + self._err())
        streams = [outs[i][: lens[i]].tobytes() if status[i] == OK else None for i in range(n)]
        return streams, recs, [status[i] for i in range(n)]

    def decode_batch(self, streams: Sequence[bytes], img_caps: Optional[Sequence[Optional[int]]] = None):
        """-> list of (img, near, effort) or None per stream; the per-stream result codes stay in self.last_status.
        img_caps[i] (optional) overrides the capacity reported for raster i (to exercise NBLIC_B200_OVERFLOW)."""
        n = len(streams)
        bufs = [np.frombuffer(s, dtype=np.uint8) if len(s) else np.zeros(1, np.uint8) for s in streams]
        lens = (C.c_size_t * max(n, 1))(*[len(s) for s in streams])
        heads = [peek(bytes(s[:16])) for s in streams]
        imgs = [np.empty(max(h[0] * h[1], 1) if h else 1, dtype=np.uint8) for h in heads]
        caps = (C.c_size_t * max(n, 1))(*[a.size if not img_caps or img_caps[i] is None else img_caps[i] for i, a in enumerate(imgs)])
        hs, ws, ns, es, status = ((C.c_int * max(n, 1))() for _ in range(5))
        rc = self.lib.nblic_b200_decode_batch(self.ctx, n, _ptr_array(bufs), lens, _ptr_array(imgs), caps, hs, ws, ns, es, status)
        if rc < 0:
            raise RuntimeError("nblic_b200_decode_batch: " + self._err())
        out = []
        self.last_status = [status[i] for i in range(n)]
        for i in range(n):
            if status[i] != OK:
                out.append(None)
            else:
                out.append((imgs[i][: hs[i] * ws[i]].reshape(hs[i], ws[i]), ns[i], es[i]))
        return out

    # ---- device-resident calls (raw CUDA pointers, e.g. torch.Tensor.data_ptr()) -------------
    def encode_device(self, d_pixels: int, pix_off: np.ndarray, heights: np.ndarray, widths: np.ndarray, near: int, effort: int,
                      d_streams: int, stream_cap: int, d_recon: int = 0):
        """-> (stream_off uint64[n+1], status int32[n], rc)"""
        n = len(heights)
        pix_off = np.ascontiguousarray(pix_off, dtype=np.uint64)
        heights = np.ascontiguousarray(heights, dtype=np.int32)
        widths = np.ascontiguousarray(widths, dtype=np.int32)
        stream_off = np.zeros(n + 1, dtype=np.uint64)
        status = np.zeros(max(n, 1), dtype=np.int32)
        rc = self.lib.nblic_b200_encode_batch_device(self.ctx, n, d_pixels, pix_off.ctypes.data_as(_u64p), heights.ctypes.data_as(_ip),
                                                     widths.ctypes.data_as(_ip), near, effort, d_streams, stream_cap,
                                                     stream_off.ctypes.data_as(_u64p), d_recon or None, status.ctypes.data_as(_ip))
        if rc < 0:
            raise RuntimeError("nblic_b200_encode_batch_device: " + self._err())
        return stream_off, status[:n], rc

    def decode_device(self, d_streams: int, stream_off: np.ndarray, d_pixels: int, pix_off: np.ndarray, pix_cap: Optional[np.ndarray] = None):
        """pix_cap[i] = bytes reserved for raster i (None: unchecked, trusted streams only) -> (status int32[n], rc)"""
        n = len(pix_off)
        stream_off = np.ascontiguousarray(stream_off, dtype=np.uint64)
        pix_off = np.ascontiguousarray(pix_off, dtype=np.uint64)
        if pix_cap is not None:
            pix_cap = np.ascontiguousarray(pix_cap, dtype=np.uint64)
        status = np.zeros(max(n, 1), dtype=np.int32)
        rc = self.lib.nblic_b200_decode_batch_device(self.ctx, n, d_streams, stream_off.ctypes.data_as(_u64p), d_pixels,
                                                     pix_off.ctypes.data_as(_u64p),
                                                     pix_cap.ctypes.data_as(_u64p) if pix_cap is not None else None,
                                                     status.ctypes.data_as(_ip))
        if rc < 0:
            raise RuntimeError("nblic_b200_decode_batch_device: " + self._err())
        return status[:n], rc

    def divcheck(self, num: np.ndarray, den: np.ndarray) -> np.ndarray:
        num = np.ascontiguousarray(num, dtype=np.int64)
        den = np.ascontiguousarray(den, dtype=np.int64)
        out = np.zeros_like(num)
        if self.lib.nblic_b200_debug_divcheck(self.ctx, num.ctypes.data, den.ctypes.data, len(num), out.ctypes.data) != 0:
            raise RuntimeError("nblic_b200_debug_divcheck: " + self._err())
        return out

    def synth_device(self, d_out: int, h: int, w: int, seed: int):
        from .synth import occluders
        occ = np.ascontiguousarray(occluders(h, w, seed), dtype=np.int32)
        rc = self.lib.nblic_b200_synth_gray(self.ctx, d_out, h, w, seed & 0xFFFFFFFF, occ.ctypes.data_as(C.POINTER(C.c_int32)))
        if rc != 0:
            raise RuntimeError("nblic_b200_synth_gray: " + self._err())


    def synth_device_batch(self, d_out: int, n: int, h: int, w: int, seed0: int, seed_stride: int = 1):
        """n images of h x w, seeds seed0 + k * seed_stride, packed back to back at d_out (one launch per 32768 images)"""
        from .synth import occluders
        occ = np.ascontiguousarray(np.stack([occluders(h, w, seed0 + i * seed_stride) for i in range(n)]), dtype=np.int32)
        rc = self.lib.nblic_b200_synth_gray_batch(self.ctx, d_out, n, h, w, seed0 & 0xFFFFFFFF, seed_stride & 0xFFFFFFFF,
                                                  occ.ctypes.data_as(C.POINTER(C.c_int32)))
        if rc != 0:
            raise RuntimeError("nblic_b200_synth_gray_batch: " + self._err())

    @property
    def last_slots(self) -> int:
        return int(self.lib.nblic_b200_last_slots(self.ctx))


class legacy:
    """The reference's own entry points, called the way src/NBLIC_main.c:184-189,223-226 calls them."""

    @staticmethod
    def nblic_compress(img: np.ndarray, near: int, effort: int):
        """-> (bytes | None, image buffer after the call, near_used, effort_used)"""
        lib = load_library()
        work = np.ascontiguousarray(img, dtype=np.uint8).copy()
        h, w = work.shape
        out = np.zeros(stream_bound(h, w), dtype=np.uint8)
        n_, e_ = C.c_int(near), C.c_int(effort)
        n = lib.NBLICcompress(0, out.ctypes.data_as(_u8p), work.ctypes.data_as(_u8p), h, w, C.byref(n_), C.byref(e_))
        return (out[:n].tobytes() if n >= 0 else None), work, n_.value, e_.value

    @staticmethod
    def nblic_decompress(data: bytes, img_capacity: int = 0):
        lib = load_library()
        head = peek(data[:16])
        cap = max(img_capacity, head[0] * head[1] if head else 0, 1)
        buf = np.zeros(max(len(data), 16) + 8, dtype=np.uint8)
        buf[: len(data)] = np.frombuffer(data, dtype=np.uint8)
        img = np.zeros(cap, dtype=np.uint8)
        hh, ww, nn, ee = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib.nblic_b200_hint_input_len(len(data))
        rc = lib.NBLICdecompress(0, buf.ctypes.data_as(_u8p), img.ctypes.data_as(_u8p), C.byref(hh), C.byref(ww), C.byref(nn), C.byref(ee))
        lib.nblic_b200_hint_input_len(0)
        if rc != 0:
            return None
        return img[: hh.value * ww.value].reshape(hh.value, ww.value), nn.value, ee.value

    @staticmethod
    def qnblic_compress(img: np.ndarray, multithread: bool = False):
        lib = load_library()
        work = np.ascontiguousarray(img, dtype=np.uint8).copy()
        h, w = work.shape
        out = np.zeros(stream_bound(h, w) // 2 + 1, dtype=np.uint16)
        fn = lib.QNBLICcompressMultiThread if multithread else lib.QNBLICcompress
        n = fn(out.ctypes.data_as(_u16p), work.ctypes.data_as(_u8p), h, w)
        return out[:n].tobytes() if n >= 0 else None

    @staticmethod
    def qnblic_decompress(data: bytes):
        lib = load_library()
        head = peek(data[:16])
        cap = max(head[0] * head[1] if head and head[3] == 0 else 0, 1)
        buf = np.zeros(max(len(data), 16) // 2 + 8, dtype=np.uint16)
        buf.view(np.uint8)[: len(data)] = np.frombuffer(data, dtype=np.uint8)
        img = np.zeros(cap, dtype=np.uint8)
        hh, ww = C.c_int(), C.c_int()
        lib.nblic_b200_hint_input_len(len(data))
        rc = lib.QNBLICdecompress(buf.ctypes.data_as(_u16p), img.ctypes.data_as(_u8p), C.byref(hh), C.byref(ww))
        lib.nblic_b200_hint_input_len(0)
        if rc != 0:
            return None
        return img[: hh.value * ww.value].reshape(hh.value, ww.value)
