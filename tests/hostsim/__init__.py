"""TEST INFRASTRUCTURE: binds the public API classes to libstarway_hostsim.so — the real
host progress engine (starway_b200/csrc/engine.cpp) linked against a CPU stand-in for the
device backend (tests/hostsim/gpu_sim.cpp).  Used only by `pytest -m "not gpu"` to exercise
connection / protocol / flush / close logic without a GPU.  Not importable from the product
package and never a fallback for it."""
import ctypes
import os

import numpy as np

from starway_b200 import _core

# SW_HOSTSIM_LIB: an instrumented build of the same library (AddressSanitizer runs of the CPU suite)
LIB_PATH = os.environ.get("SW_HOSTSIM_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libstarway_hostsim.so")


def load(use_fastpath: bool = True):
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} missing: run `make hostsim`")
    return _core.bind(ctypes.CDLL(LIB_PATH), lambda: int(os.environ.get("SW_SIM_DEVICE", "0")), use_fastpath=use_fastpath)


class SimDev:
    """'Device' buffers of the CPU stand-in backend: memory from its device allocator (which the engine's
    pointer query reports as device memory), exposed to NumPy for the test and to the binding through
    `__cuda_array_interface__` — the engine then takes its device-buffer paths (eager straight from the
    buffer, zero-copy rendezvous, IPC export of the allocation)."""

    _lib = None

    class Buf:
        def __init__(self, lib, nbytes, exportable=True):
            self.nbytes = int(nbytes)
            self._lib = lib
            self.ptr = (lib.swsim_dev_alloc if exportable else lib.swsim_dev_alloc_noexport)(max(self.nbytes, 1))
            if not self.ptr:
                raise MemoryError("swsim_dev_alloc failed")
            self.np = np.ctypeslib.as_array((ctypes.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr))[: self.nbytes]

        @property
        def __cuda_array_interface__(self):
            return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}

        def __del__(self):
            try:
                if self.ptr:
                    self.np = None
                    self._lib.swsim_dev_free(ctypes.c_void_p(self.ptr))
                    self.ptr = 0
            except Exception:
                pass

    @classmethod
    def lib(cls):
        if cls._lib is None:
            lib = ctypes.CDLL(LIB_PATH)
            lib.swsim_dev_alloc.restype = ctypes.c_void_p
            lib.swsim_dev_alloc.argtypes = [ctypes.c_size_t]
            lib.swsim_dev_alloc_noexport.restype = ctypes.c_void_p
            lib.swsim_dev_alloc_noexport.argtypes = [ctypes.c_size_t]
            lib.swsim_dev_free.argtypes = [ctypes.c_void_p]
            cls._lib = lib
        return cls._lib

    # the buffer factory protocol of tests/cases_basic.py (HostBufs)
    @classmethod
    def alloc(cls, cap):
        b = cls.Buf(cls.lib(), cap)
        b.np[:] = 0xEE
        return b

    @classmethod
    def from_np(cls, a, exportable=True):
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
        b = cls.Buf(cls.lib(), a.nbytes, exportable)
        b.np[:] = a
        return b

    @staticmethod
    def to_np(b):
        return np.array(b.np, copy=True)

    @staticmethod
    def sync():
        return None
