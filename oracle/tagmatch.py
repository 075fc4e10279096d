"""oracle/tagmatch.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two views of the same CPU restatement of the UCP tag-matching contract the reference
relies on (libucp behind reference src/bindings/main.cpp:370,404,1136,1172; see
oracle/tagmatch.h for provenance and what is / is not pinned by the reference's tests):

  * ``COracle``  — ctypes wrapper over oracle/liboracle_tagmatch.so (oracle/tagmatch.c);
  * ``PyOracle`` — an independent pure-Python mirror (lists + loops) for small cases,
    used to cross-check the C restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import this module.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

ORC_OK = 0
ORC_ERR_MESSAGE_TRUNCATED = -9
ORC_ERR_CANCELED = -16

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_tagmatch.so")
U64 = (1 << 64) - 1


def tag_match(stag: int, tag: int, mask: int) -> bool:
    """recv (tag, mask) accepts sender tag stag iff ((stag ^ tag) & mask) == 0."""
    return ((stag ^ tag) & mask & U64) == 0


@dataclass
class Match:
    op_id: int
    sender_tag: int
    length: int
    status: int
    ep: int = 0
    user: int = 0


class _OrcMatch(ctypes.Structure):
    _fields_ = [
        ("op_id", ctypes.c_uint64),
        ("sender_tag", ctypes.c_uint64),
        ("length", ctypes.c_uint64),
        ("status", ctypes.c_int32),
        ("ep", ctypes.c_uint32),
        ("user", ctypes.c_uint64),
    ]


def _load():
    lib = ctypes.CDLL(LIB_PATH)
    lib.orc_worker_new.restype = ctypes.c_void_p
    lib.orc_worker_free.argtypes = [ctypes.c_void_p]
    lib.orc_post_recv.restype = ctypes.c_int
    lib.orc_post_recv.argtypes = [
        ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
        ctypes.POINTER(_OrcMatch),
    ]
    lib.orc_arrive.restype = ctypes.c_int
    lib.orc_arrive.argtypes = [
        ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64,
        ctypes.POINTER(_OrcMatch),
    ]
    lib.orc_cancel_all.restype = ctypes.c_size_t
    lib.orc_cancel_all.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.c_size_t]
    lib.orc_num_posted.restype = ctypes.c_size_t
    lib.orc_num_posted.argtypes = [ctypes.c_void_p]
    lib.orc_num_unexpected.restype = ctypes.c_size_t
    lib.orc_num_unexpected.argtypes = [ctypes.c_void_p]
    lib.orc_tag_match.restype = ctypes.c_int
    lib.orc_tag_match.argtypes = [ctypes.c_uint64] * 3
    return lib


class COracle:
    """One matching domain (= one reference Client or Server object)."""

    _lib = None

    def __init__(self):
        if COracle._lib is None:
            COracle._lib = _load()
        self._w = COracle._lib.orc_worker_new()
        self._bufs: dict[int, np.ndarray] = {}

    def __del__(self):
        try:
            COracle._lib.orc_worker_free(self._w)
        except Exception:
            pass

    @staticmethod
    def _m(m: _OrcMatch) -> Match:
        return Match(m.op_id, m.sender_tag, m.length, m.status, m.ep, m.user)

    def post_recv(self, op_id: int, tag: int, mask: int, buf: np.ndarray) -> Match | None:
        assert buf.dtype == np.uint8 and buf.flags.c_contiguous
        self._bufs[op_id] = buf  # keep alive while posted
        m = _OrcMatch()
        r = COracle._lib.orc_post_recv(self._w, op_id, tag & U64, mask & U64, buf.ctypes.data, buf.nbytes, ctypes.byref(m))
        if r:
            self._bufs.pop(op_id, None)
            return self._m(m)
        return None

    def arrive(self, ep: int, stag: int, data: np.ndarray | None, length: int | None = None, user: int = 0):
        m = _OrcMatch()
        if data is None:
            r = COracle._lib.orc_arrive(self._w, ep, stag & U64, None, length or 0, user, ctypes.byref(m))
        else:
            data = np.ascontiguousarray(data, dtype=np.uint8)
            r = COracle._lib.orc_arrive(self._w, ep, stag & U64, data.ctypes.data, data.nbytes, user, ctypes.byref(m))
        if r:
            self._bufs.pop(m.op_id, None)
            return self._m(m)
        return None

    def cancel_all(self) -> list[int]:
        arr = (ctypes.c_uint64 * 65536)()
        n = COracle._lib.orc_cancel_all(self._w, arr, 65536)
        self._bufs.clear()
        return [arr[i] for i in range(min(n, 65536))]

    @property
    def num_posted(self) -> int:
        return COracle._lib.orc_num_posted(self._w)

    @property
    def num_unexpected(self) -> int:
        return COracle._lib.orc_num_unexpected(self._w)


class PyOracle:
    """Pure-Python mirror of oracle/tagmatch.c (small cases only)."""

    def __init__(self):
        self.posted: list[tuple[int, int, int, np.ndarray]] = []  # (op, tag, mask, buf) in post order
        self.unexp: list[tuple[int, int, bytes | None, int, int]] = []  # (ep, stag, data, len, user) in arrival order

    @staticmethod
    def _deliver(op, stag, data, length, ep, user, buf) -> Match:
        if length > buf.nbytes:
            return Match(op, stag, length, ORC_ERR_MESSAGE_TRUNCATED, ep, user)
        if data is not None and length:
            buf[:length] = np.frombuffer(data, dtype=np.uint8)
        return Match(op, stag, length, ORC_OK, ep, user)

    def post_recv(self, op_id, tag, mask, buf):
        for i, (ep, stag, data, length, user) in enumerate(self.unexp):
            if tag_match(stag, tag, mask):
                del self.unexp[i]
                return self._deliver(op_id, stag, data, length, ep, user, buf)
        self.posted.append((op_id, tag, mask, buf))
        return None

    def arrive(self, ep, stag, data, length=None, user=0):
        payload = None if data is None else bytes(np.ascontiguousarray(data, dtype=np.uint8))
        n = (length or 0) if data is None else len(payload)
        for i, (op, tag, mask, buf) in enumerate(self.posted):
            if tag_match(stag, tag, mask):
                del self.posted[i]
                return self._deliver(op, stag, payload, n, ep, user, buf)
        self.unexp.append((ep, stag, payload, n, user))
        return None

    def cancel_all(self):
        ops = [p[0] for p in self.posted]
        self.posted.clear()
        return ops

    @property
    def num_posted(self):
        return len(self.posted)

    @property
    def num_unexpected(self):
        return len(self.unexp)
