"""CPU tests of the oracle (oracle/tagmatch.c + its pure-Python mirror):
  * against the golden vectors transcribed from the reference's own tests
    (tests/golden/reference_cases.json, made by tests/golden/make_golden.py);
  * C restatement vs independent Python mirror on random schedules;
  * the documented UCX rules no reference test pins (partial masks, truncation, order)."""
import numpy as np
import pytest

from oracle.tagmatch import ORC_ERR_MESSAGE_TRUNCATED, ORC_OK, COracle, PyOracle, tag_match
from tests.golden_util import load_cases, payload

CASES = load_cases()


def run_schedule(orc, events):
    bufs, done = {}, {}
    for ev in events:
        if ev[0] == "recv":
            _, op, tag, mask, cap = ev
            bufs[op] = np.full(cap, 0xEE, dtype=np.uint8)
            m = orc.post_recv(op, tag, mask, bufs[op])
        else:
            _, ep, tag, spec = ev
            m = orc.arrive(ep, tag, payload(spec))
        if m is not None:
            done[m.op_id] = m
    return bufs, done


@pytest.mark.parametrize("impl", [COracle, PyOracle], ids=["c", "py"])
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_reference_cases(impl, case):
    orc = impl()
    bufs, done = run_schedule(orc, case["events"])
    for op, (tag, length, spec) in case.get("complete", {}).items():
        m = done[int(op)]
        assert (m.sender_tag, m.length, m.status) == (tag, length, ORC_OK)
        if spec is not None:
            np.testing.assert_array_equal(bufs[int(op)][:length], payload(spec))
            assert (bufs[int(op)][length:] == 0xEE).all()  # bytes beyond `length` untouched
    if "tagset" in case:
        ops, tags = case["tagset"]
        assert {done[o].sender_tag for o in ops} == set(tags)
        assert all(done[o].status == ORC_OK for o in ops)
    for op in case.get("pending", []):
        assert op not in done
        assert op in orc.cancel_all()  # close => "Request canceled"


def test_match_rule_masks():
    # SURVEY Appendix A.2; partial masks are pinned only by the documented UCX rule
    assert tag_match(0x1234, 0xFF34, 0x00FF)
    assert not tag_match(0x1235, 0xFF34, 0x00FF)
    assert tag_match(999, 5, 0)  # mask 0 = wildcard regardless of tag (reference tests :650)
    assert tag_match((1 << 64) - 1, (1 << 64) - 1, (1 << 64) - 1)
    assert not tag_match(1 << 63, 0, 1 << 63)


@pytest.mark.parametrize("impl", [COracle, PyOracle], ids=["c", "py"])
def test_order_truncation_zero_length(impl):
    orc = impl()
    # earliest-posted matching receive wins
    b1, b2, b3 = (np.zeros(8, dtype=np.uint8) for _ in range(3))
    assert orc.post_recv(1, 7, 0xFF, b1) is None
    assert orc.post_recv(2, 0, 0, b2) is None
    assert orc.post_recv(3, 7, 0xFF, b3) is None
    m = orc.arrive(0, 0x107, np.array([1, 2, 3], dtype=np.uint8))
    assert (m.op_id, m.sender_tag, m.length, m.status) == (1, 0x107, 3, ORC_OK)
    m = orc.arrive(0, 0x999, np.array([9], dtype=np.uint8))
    assert m.op_id == 2  # wildcard takes anything
    # earliest-arrived matching unexpected message wins
    for i, t in enumerate([5, 6, 5]):
        assert orc.arrive(1, t, np.array([i], dtype=np.uint8)) is None
    rb = np.zeros(4, dtype=np.uint8)
    m = orc.post_recv(10, 5, 0xFFFF, rb)
    assert (m.sender_tag, m.length, rb[0]) == (5, 1, 0)
    m = orc.post_recv(11, 5, 0xFFFF, rb)
    assert (m.sender_tag, rb[0]) == (5, 2)
    # truncation: message consumed, buffer untouched
    small = np.full(2, 0xEE, dtype=np.uint8)
    assert orc.post_recv(20, 77, (1 << 64) - 1, small) is None
    m = orc.arrive(0, 77, np.arange(5, dtype=np.uint8))
    assert (m.op_id, m.length, m.status) == (20, 5, ORC_ERR_MESSAGE_TRUNCATED)
    assert (small == 0xEE).all()
    # zero-length messages are legal
    z = np.full(3, 0xEE, dtype=np.uint8)
    # op 3 (tag 7, mask 0xFF) and the unexpected tag-6 message are still queued
    assert orc.post_recv(30, 1, 0xF, z) is None
    m = orc.arrive(2, 0x1231, np.zeros(0, dtype=np.uint8))
    assert (m.op_id, m.length, m.status) == (30, 0, ORC_OK) and (z == 0xEE).all()
    assert orc.num_posted == 1 and orc.num_unexpected == 1
    assert orc.cancel_all() == [3]


def test_c_vs_python_random_schedules():
    rng = np.random.default_rng(20260921)
    masks = [0, (1 << 64) - 1, 0xFF, 0xF0, 0xFFFF]
    for trial in range(200):
        c, p = COracle(), PyOracle()
        cb, pb = {}, {}
        op = 1
        for _ in range(int(rng.integers(10, 120))):
            if rng.random() < 0.5:
                tag, mask = int(rng.integers(0, 6)), masks[int(rng.integers(0, 5))]
                cap = int(rng.choice([0, 1, 4, 16, 64]))
                cb[op], pb[op] = np.full(cap, 0xEE, np.uint8), np.full(cap, 0xEE, np.uint8)
                mc, mp_ = c.post_recv(op, tag, mask, cb[op]), p.post_recv(op, tag, mask, pb[op])
                op += 1
            else:
                ep, stag = int(rng.integers(0, 3)), int(rng.integers(0, 6)) | (int(rng.integers(0, 2)) << 8)
                data = rng.integers(0, 256, int(rng.choice([0, 1, 3, 16, 40])), dtype=np.uint8)
                mc, mp_ = c.arrive(ep, stag, data), p.arrive(ep, stag, data)
            assert (mc is None) == (mp_ is None)
            if mc is not None:
                assert (mc.op_id, mc.sender_tag, mc.length, mc.status, mc.ep) == (
                    mp_.op_id, mp_.sender_tag, mp_.length, mp_.status, mp_.ep)
        for k in cb:
            np.testing.assert_array_equal(cb[k], pb[k])
        assert (c.num_posted, c.num_unexpected) == (p.num_posted, p.num_unexpected)
        assert sorted(c.cancel_all()) == sorted(p.cancel_all())
