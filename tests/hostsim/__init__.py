"""TEST INFRASTRUCTURE: binds the public API classes to libstarway_hostsim.so — the real
host progress engine (starway_b200/csrc/engine.cpp) linked against a CPU stand-in for the
device backend (tests/hostsim/gpu_sim.cpp).  Used only by `pytest -m "not gpu"` to exercise
connection / protocol / flush / close logic without a GPU.  Not importable from the product
package and never a fallback for it."""
import ctypes
import os

from starway_b200 import _core

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libstarway_hostsim.so")


def load(use_fastpath: bool = True):
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} missing: run `make hostsim`")
    return _core.bind(ctypes.CDLL(LIB_PATH), lambda: int(os.environ.get("SW_SIM_DEVICE", "0")), use_fastpath=use_fastpath)
