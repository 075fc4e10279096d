"""Builds tests/golden/reference_cases.json — the outcomes the REFERENCE's own tests assert
for the tagged send/recv path, restated as event schedules for one matching domain.

Provenance: /root/reference/tests/test_basic.py (cited per case).  The reference itself
cannot be imported or run here (its `_bindings` extension needs nanobind + libucp/UCX
1.18.1, both absent from this image and from the GPU box), so these
vectors are transcribed from the assertions in its test file rather than recorded from a
run.  Payloads are given symbolically ("arange:N", "fill:V:N", "seed:S:N") and expanded
by tests/test_oracle.py.

Event forms:
  ["recv", op_id, tag, mask, cap]          receive posted on the worker under test
  ["send", ep, tag, payload]               message arriving from endpoint `ep`
Expectations:
  "complete": {op_id: [sender_tag, length, payload-or-null]}  (exact, per receive)
  "tagset":   [ops, tags]  the set of sender tags returned by `ops` equals `tags`
  "pending":  [op_id, ...] receives that must NOT have completed (cancelled at close)

Run:  python tests/golden/make_golden.py
"""
import json
import os

U64 = (1 << 64) - 1
cases = []

# tests/test_basic.py:146-163  client -> server, recv posted first with (tag=0, mask=0)
cases.append({
    "name": "client_to_server_send_recv", "ref": "tests/test_basic.py:146-163",
    "events": [["recv", 1, 0, 0, 10], ["send", 0, 1, "arange:10"]],
    "complete": {"1": [1, 10, "arange:10"]},
})
# tests/test_basic.py:166-187  server -> client
cases.append({
    "name": "server_to_client_send_recv", "ref": "tests/test_basic.py:166-187",
    "events": [["recv", 1, 0, 0, 20], ["send", 0, 2, "arange:20"]],
    "complete": {"1": [2, 20, "arange:20"]},
})
# tests/test_basic.py:61-101  address mode, 16 B both ways, tags 1 and 2 (each direction = its own domain)
cases.append({
    "name": "worker_address_roundtrip_s2c", "ref": "tests/test_basic.py:79-87",
    "events": [["recv", 1, 0, 0, 16], ["send", 0, 1, "arange:16"]],
    "complete": {"1": [1, 16, "arange:16"]},
})
cases.append({
    "name": "worker_address_roundtrip_c2s", "ref": "tests/test_basic.py:89-96",
    "events": [["recv", 1, 0, 0, 16], ["send", 0, 2, "arange:16"]],
    "complete": {"1": [2, 16, "arange:16"]},
})
# tests/test_basic.py:418-442  sizes 1/1024/4096, random bytes, tags 3 (c->s) and 4 (s->c)
for size in (1, 1024, 4096):
    for tag in (3, 4):
        cases.append({
            "name": f"message_integrity_{size}_tag{tag}", "ref": "tests/test_basic.py:418-442",
            "events": [["recv", 1, 0, 0, size], ["send", 0, tag, f"seed:{size + tag}:{size}"]],
            "complete": {"1": [tag, size, f"seed:{size + tag}:{size}"]},
        })
# tests/test_basic.py:526-554  unexpected queue: 5 clients send 1 B tag=i BEFORE any receive is posted
ev = [["send", i, i, f"fill:{i}:1"] for i in range(5)] + [["recv", 10 + i, 0, 0, 1] for i in range(5)]
cases.append({
    "name": "multiple_clients_unexpected", "ref": "tests/test_basic.py:526-554",
    "events": ev, "tagset": [[10, 11, 12, 13, 14], [0, 1, 2, 3, 4]],
})
# tests/test_basic.py:557-570  50 sends (1 byte each, value i & 0xFF) and 50 wildcard receives
ev = [["send", 0, i, f"fill:{i & 255}:1"] for i in range(50)] + [["recv", 100 + i, 0, 0, 1] for i in range(50)]
cases.append({
    "name": "concurrent_send_recv", "ref": "tests/test_basic.py:557-570",
    "events": ev, "tagset": [[100 + i for i in range(50)], list(range(50))],
})
# tests/test_basic.py:573-610  2000 messages each way; one direction shown per domain, receives posted first
for base in (100, 200):
    ev = [["recv", 1000 + i, 0, 0, 1] for i in range(2000)] + [["send", 0, base + i, f"fill:{i & 255}:1"] for i in range(2000)]
    cases.append({
        "name": f"bidirectional_traffic_{base}", "ref": "tests/test_basic.py:573-610",
        "events": ev, "tagset": [[1000 + i for i in range(2000)], [base + i for i in range(2000)]],
    })
# tests/test_basic.py:613-630  10 clients each connect/send(tag 1)/close vs 10 posted receives
ev = [["recv", 1 + i, 0, 0, 1] for i in range(10)] + [["send", i, 1, "fill:0:1"] for i in range(10)]
cases.append({
    "name": "rapid_connect_close", "ref": "tests/test_basic.py:613-630",
    "events": ev, "tagset": [[1 + i for i in range(10)], [1]],
})
# tests/test_basic.py:638-663  arecv(buf, 999, 0) never matched -> cancelled at close ("cancel" in message)
cases.append({
    "name": "shutdown_with_in_flight_ops", "ref": "tests/test_basic.py:638-663",
    "events": [["recv", 1, 999, 0, 1024]], "pending": [1],
})
# README.md:60-67 / BASELINE config 1+2: tag=1, tag_mask=0xFFFF
cases.append({
    "name": "readme_quickstart_mask_ffff", "ref": "README.md:60-67",
    "events": [["recv", 1, 1, 0xFFFF, 4], ["send", 0, 1, "arange:4"]],
    "complete": {"1": [1, 4, "arange:4"]},
})
# benchmarks/scenarios.py:10  full-mask receives (TAG_MASK = 2^64-1) used by the reference bench
cases.append({
    "name": "bench_full_mask", "ref": "src/starway/benchmarks/scenarios.py:10-22",
    "events": [["recv", 1, 0x2B00, U64, 64], ["send", 0, 0x2B10, "fill:7:8"], ["send", 0, 0x2B00, "fill:9:64"]],
    "complete": {"1": [0x2B00, 64, "fill:9:64"]},
})

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_cases.json")
with open(out, "w") as f:
    json.dump(cases, f, separators=(",", ":"))
print(f"wrote {len(cases)} cases to {out} ({os.path.getsize(out)} bytes)")
