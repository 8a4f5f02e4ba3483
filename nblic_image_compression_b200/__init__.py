"""nblic_image_compression_b200 -- B200 (sm_100a) implementation of NBLIC's encode/decode hot path.

The product is the shared library `libnblic_b200.so` (CUDA kernels + C ABI, see include/nblic_b200.h);
this package is the thin host-side mirror of the reference interface used by the tests and bench.py:

    api.Codec            batch encode / decode through the C ABI (host buffers or device pointers)
    api.legacy           the reference's five entry points (NBLICcompress, ...) through the same library
    synth.gen            the deterministic synthetic test images (numpy definition)
    build.build_library  nvcc build of the library, in-tree

There is no CPU fallback: loading fails loudly when the library is missing, and creating a Codec
fails when no CUDA device is usable.
"""
from .build import LIB, build_library  # noqa: F401

__all__ = ["LIB", "build_library"]
