"""Device plumbing over the C ABI: handle, device buffers, pinned host buffers."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class Handle:
    """One bbm_handle_t (device + stream).  Not thread-safe; create one per thread/stream."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = L.load()
        h = C.c_void_p()
        rc = self._lib.bbm_create(device, C.byref(h))
        if rc != 0:
            raise L.BBMError(f"bbm_create(device={device}) failed: {self._lib.bbm_status_string(rc).decode()} "
                             "(a CUDA device of compute capability 10.x is required; there is no CPU fallback)")
        self.h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, stream: int):
        L.check(self._lib.bbm_set_stream(self.h, C.c_void_p(stream)), self.h)

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        L.check(self._lib.bbm_synchronize(self.h), self.h)

    def set_option(self, key: str, value: int):
        L.check(self._lib.bbm_set_option(self.h, key.encode(), int(value)), self.h)

    def get_option(self, key: str) -> int:
        v = C.c_int64()
        L.check(self._lib.bbm_get_option(self.h, key.encode(), C.byref(v)), self.h)
        return v.value

    @property
    def launch_count(self) -> int:
        v = C.c_int64()
        L.check(self._lib.bbm_launch_count(self.h, C.byref(v)), self.h)
        return v.value

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self._lib.bbm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- buffers -------------------------------------------------------------------------
    def empty(self, shape, dtype) -> "DeviceArray":
        return DeviceArray(self, shape, dtype)

    def to_device(self, a: np.ndarray) -> "DeviceArray":
        d = DeviceArray(self, a.shape, a.dtype)
        d.copy_from_host(a)
        return d

    def pinned(self, shape, dtype) -> "PinnedArray":
        return PinnedArray(self, shape, dtype)


class DeviceArray:
    """A column-major device buffer allocated with bbm_malloc (owned) or wrapping a raw pointer."""

    def __init__(self, handle: Handle, shape, dtype, ptr: int | None = None):
        self.handle = handle
        self.shape = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._owned = ptr is None
        if ptr is None:
            p = C.c_void_p()
            L.check(handle._lib.bbm_malloc(handle.h, C.byref(p), self.nbytes), handle.h, "bbm_malloc")
            ptr = p.value
        self.ptr = int(ptr)

    @property
    def ld(self) -> int:
        return max(1, self.shape[0])

    def copy_from_host(self, a: np.ndarray):
        a = np.asarray(a, dtype=self.dtype)
        if tuple(a.shape) != self.shape:
            raise L.DimensionMismatch(f"host array {a.shape} vs device array {self.shape}")
        a = np.asfortranarray(a)
        h = self.handle
        L.check(h._lib.bbm_memcpy_h2d(h.h, C.c_void_p(self.ptr), a.ctypes.data_as(C.c_void_p), self.nbytes), h.h)
        h.synchronize()  # pageable source: do not let `a` die before the copy ran
        return self

    def to_host(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype, order="F")
        h = self.handle
        L.check(h._lib.bbm_memcpy_d2h(h.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), self.nbytes), h.h)
        h.synchronize()
        return out

    def fill_bytes(self, byte: int):
        h = self.handle
        L.check(h._lib.bbm_memset(h.h, C.c_void_p(self.ptr), byte, self.nbytes), h.h)
        return self

    def free(self):
        if self._owned and self.ptr and self.handle.h:
            self.handle._lib.bbm_free(self.handle.h, C.c_void_p(self.ptr))
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """Page-locked host memory from bbm_host_alloc, exposed as a column-major numpy array."""

    def __init__(self, handle: Handle, shape, dtype):
        self.handle = handle
        shape = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
        dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64))
        p = C.c_void_p()
        L.check(handle._lib.bbm_host_alloc(handle.h, C.byref(p), n * dtype.itemsize), handle.h, "bbm_host_alloc")
        self.ptr = p.value
        buf = (C.c_char * max(1, n * dtype.itemsize)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape, order="F")

    def free(self):
        if self.ptr and self.handle.h:
            self.array = None
            self.handle._lib.bbm_host_free(self.handle.h, C.c_void_p(self.ptr))
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
