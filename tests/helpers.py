"""Shared helpers for the test-suite: golden vectors (minted from the compiled reference by
oracle/make_golden.py) and small utilities."""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "golden")

_cache = {}


def golden():
    if "g" not in _cache:
        _cache["g"] = json.load(open(os.path.join(GOLDEN_DIR, "golden.json")))
        _cache["ev"] = np.load(os.path.join(GOLDEN_DIR, "events.npz"))
    return _cache["g"], _cache["ev"]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs], dtype=np.float64)


def case_events(case):
    _, ev = golden()
    k = case["events"]
    fx, fy, t = ev[k + "_fr_x"], ev[k + "_fr_y"], ev[k + "_t_ns"]
    n = case.get("first_n")
    if n:
        fx, fy, t = fx[:n], fy[:n], t[:n]
    noise = None
    if case.get("noise_every"):
        noise = (np.arange(len(fx)) % case["noise_every"] == 0).astype(np.uint8)
    init = unhex(case["init"]) if case.get("init") else None
    return fx, fy, t, noise, init


def stage_positions(fr_x, fr_y):
    """Same deterministic jitter as oracle/make_golden.py:stage_positions."""
    i = np.arange(len(fr_x), dtype=np.int64)
    pr_x = fr_x.astype(np.float64) + ((i * 7919) % 1000 - 500).astype(np.float64) / 256.0
    pr_y = fr_y.astype(np.float64) + ((i * 104729) % 1000 - 500).astype(np.float64) / 256.0
    return pr_x, pr_y


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def same_model(a, b, rtol=1e-12):
    """Two results of the same slice from launches with a different CTA grouping (batched vs alone, helped vs
    not, another device count).  The integer image sums are order-independent, so occupancy counts must be
    equal; the fp64 gradient moments are summed in a grouping-dependent order and may differ in the last bits."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return a[6] == b[6] and bool(np.allclose(a, b, rtol=rtol, atol=0))


def ring_slice(ts, consumed, capacity=50000, span_ns=200_000_000):
    """The events DVS_flow<capacity, span>::recompute hands to the optimiser after `consumed` events of a stream
    with timestamps `ts` (ns, non-decreasing): indices newest -> oldest and the slice start time -- the ring buffer
    with its lazy eviction and its full-buffer quirk restated (datastructures.h:31-96, dvs_flow.h:186-198; the same
    restatement tests/test_slicing_cpu.py pins against the compiled reference)."""
    ts = np.asarray(ts).astype(np.int64)
    c = int(consumed)
    newest = int(ts[c - 1])
    lo = int(np.searchsorted(ts[:c], newest - span_ns, side="right")) if newest >= span_ns else 0
    lo = max(lo, c - capacity)
    full = (c - lo) == capacity
    first = lo + (1 if full else 0)
    idx = np.arange(c - 1, first - 1, -1)
    start = int(ts[lo]) if full else (newest - span_ns if newest > span_ns else 0)
    return idx, start
