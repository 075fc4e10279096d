"""Helpers shared by the oracle tests and the GPU parity tests: golden-vector loading and
symbolic payload expansion (see tests/golden/make_golden.py)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cases.json")


def load_cases():
    with open(GOLDEN) as f:
        return json.load(f)


def payload(spec: str) -> np.ndarray:
    kind, *rest = spec.split(":")
    if kind == "arange":
        return (np.arange(int(rest[0])) & 0xFF).astype(np.uint8)
    if kind == "fill":
        return np.full(int(rest[1]), int(rest[0]), dtype=np.uint8)
    if kind == "seed":
        return np.random.default_rng(int(rest[0])).integers(0, 256, int(rest[1]), dtype=np.uint8)
    raise ValueError(spec)
