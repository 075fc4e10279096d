"""oracle/starway_cpu.py — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

asyncio front-end over oracle/libstarway_cpu.so (oracle/cpu_engine.cpp): the "restated
reference" whose execution shape follows the reference package
(reference src/starway/__init__.py:71-345 + src/bindings/main.cpp): every operation creates
an asyncio Future and two closures, the native worker thread calls the closure under the GIL,
and the closure bounces to the loop with call_soon_threadsafe — one wake-up per operation.
Host (NumPy) buffers only, Server and Client in the same process.  Used solely as the timed
CPU baseline by bench.py and by tests/test_cpu_baseline.py.
"""
from __future__ import annotations

import asyncio
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstarway_cpu.so")

DONE = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64)
FAIL = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_char_p)

_lib = None


def lib():
    global _lib
    if _lib is None:
        l = ctypes.CDLL(LIB_PATH)
        l.swc_worker_new.restype = ctypes.c_void_p
        l.swc_worker_free.argtypes = [ctypes.c_void_p]
        l.swc_listen.argtypes = [ctypes.c_void_p, ctypes.c_int]
        l.swc_connect.argtypes = [ctypes.c_void_p, ctypes.c_int]
        l.swc_send.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, DONE, FAIL, ctypes.c_void_p]
        l.swc_recv.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint64, DONE, FAIL, ctypes.c_void_p]
        l.swc_flush.argtypes = [ctypes.c_void_p, DONE, FAIL, ctypes.c_void_p]
        l.swc_close.argtypes = [ctypes.c_void_p, DONE, ctypes.c_void_p]
        l.swc_status.argtypes = [ctypes.c_void_p]
        _lib = l
    return _lib


class _Worker:
    def __init__(self):
        self._w = lib().swc_worker_new()
        self._live = {}  # keeps callbacks + buffers alive until completion
        self._n = 0

    def __del__(self):
        try:
            lib().swc_worker_free(self._w)
        except Exception:
            pass

    def _op(self, loop, result_of):
        if loop is None:
            loop = asyncio.get_running_loop()
        fut = loop.create_future()
        self._n += 1
        key = self._n

        # the ctypes thunks are released on the loop thread, never from inside their own call
        def finish(setter, value):
            self._live.pop(key, None)
            if not fut.done():
                setter(value)

        def on_done(_user, tag, length):
            loop.call_soon_threadsafe(finish, fut.set_result, result_of(tag, length))

        def on_fail(_user, reason):
            loop.call_soon_threadsafe(finish, fut.set_exception, Exception(reason.decode()))

        d, f = DONE(on_done), FAIL(on_fail)
        self._live[key] = (d, f)
        return fut, d, f, key

    def arecv(self, buffer: np.ndarray, tag: int, tag_mask: int, loop=None):
        fut, d, f, key = self._op(loop, lambda t, n: (t, n))
        self._live[key] += (buffer,)
        if lib().swc_recv(self._w, buffer.ctypes.data, buffer.nbytes, tag, tag_mask, d, f, None) != 0:
            raise RuntimeError("not running")
        return fut

    def aflush(self, loop=None):
        fut, d, f, _ = self._op(loop, lambda t, n: None)
        if lib().swc_flush(self._w, d, f, None) != 0:
            raise RuntimeError("not running")
        return fut

    def aclose(self, loop=None):
        fut, d, f, _ = self._op(loop, lambda t, n: None)
        if lib().swc_close(self._w, d, None) != 0:
            raise RuntimeError("not running")
        return fut

    def _asend(self, ep: int, buffer: np.ndarray, tag: int, loop=None):
        if buffer.dtype != np.uint8:
            buffer = buffer.astype(np.uint8)
        fut, d, f, key = self._op(loop, lambda t, n: None)
        self._live[key] += (buffer,)
        if lib().swc_send(self._w, ep, buffer.ctypes.data, buffer.nbytes, tag, d, f, None) != 0:
            raise RuntimeError("not running")
        return fut


class Server(_Worker):
    def __init__(self):
        super().__init__()
        self._n_eps = 0

    def listen(self, addr: str, port: int):
        if lib().swc_listen(self._w, port) != 0:
            raise RuntimeError("Server: already listening.")

    def list_clients(self):
        return set(range(self._n_eps))

    def asend(self, client_ep: int, buffer, tag: int, loop=None):
        return self._asend(client_ep, buffer, tag, loop)

    def aflush_ep(self, client_ep: int, loop=None):
        return self.aflush(loop)


_servers: dict[int, Server] = {}


class Client(_Worker):
    async def aconnect(self, addr: str, port: int):
        r = lib().swc_connect(self._w, port)
        if r < 0:
            raise Exception("Endpoint is not connected")
        srv = _servers.get(port)
        if srv is not None:
            srv._n_eps = max(srv._n_eps, r + 1)

    def asend(self, buffer, tag: int, loop=None):
        return self._asend(0, buffer, tag, loop)


def make_pair(port: int):
    """Server + connected Client in this process (the reference's gen_server_client fixture shape)."""
    s, c = Server(), Client()
    s.listen("127.0.0.1", port)
    _servers[port] = s
    return s, c
