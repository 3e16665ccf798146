"""What a slice IS is decided on the host, before any kernel runs: DVS_flow::add_event's triggers, the
CircularArray ring buffer with its lazy eviction and its quirks (first push lands at index 1; a full buffer is
iterated without its oldest element, datastructures.h:35-42,73-75), the slice start time (dvs_flow.h:186-193)
and Event::set_local_time.  The host mirror (better_flow_b200/include/better_flow/dvs_flow.h) is run here in
its queue-only mode -- no GPU involved -- against the golden streams and against the compiled reference."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from better_flow_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def sl():
    so = os.path.join(HERE, "cpu", "libslicer_shim.so")
    src = os.path.join(HERE, "cpu", "slicer_shim.cpp")
    inc = os.path.join(ROOT, "better_flow_b200", "include")
    libdir = os.path.join(ROOT, "better_flow_b200")
    if not os.path.exists(os.path.join(libdir, "libbf_cuda.so")):
        subprocess.check_call(["make", "-s", "-C", libdir, "lib"])
    hdrs = [os.path.join(inc, "better_flow", f) for f in os.listdir(os.path.join(inc, "better_flow"))]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-I" + inc,
                               "-I" + os.path.join(ROOT, "include"), src, "-L" + libdir, "-lbf_cuda",
                               "-Wl,-rpath," + libdir, "-o", so])
    return C.CDLL(so)


def slices_of(lib, fr_x, fr_y, ts, config=0, ev_refresh=20000, time_refresh_ns=33000000, flush=True, capacity=0, span_ns=0,
              max_slices=4096):
    fx = np.ascontiguousarray(fr_x, dtype=np.uint32)
    fy = np.ascontiguousarray(fr_y, dtype=np.uint32)
    t = np.ascontiguousarray(ts, dtype=np.uint64)
    info = np.zeros((max_slices, 3), dtype=np.int64)
    rec = np.zeros((max_slices, 6), dtype=np.int64)
    p = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
    k = lib.sl_stream(config, len(fx), p(fx, C.c_uint32), p(fy, C.c_uint32), p(t, C.c_uint64), C.c_ulonglong(ev_refresh),
                      C.c_ulonglong(time_refresh_ns), 1 if flush else 0, max_slices, C.c_longlong(capacity),
                      C.c_longlong(span_ns), p(info, C.c_longlong), p(rec, C.c_longlong))
    assert 0 <= k <= max_slices
    return info[:k], rec[:k]


def expected_slices(fr_x, ts, consumed, capacity, span_ns):
    """The ring buffer restated: after `consumed` events the buffer holds the newest events whose age relative
    to the newest one is < span, at most `capacity` of them; when it is full the iteration skips the oldest."""
    out = []
    for c in consumed:
        newest = int(ts[c - 1])
        lo = int(np.searchsorted(ts[:c], newest - span_ns, side="right")) if newest >= span_ns else 0
        lo = max(lo, c - capacity)
        held = c - lo
        full = held == capacity
        first = lo + (1 if full else 0)                    # oldest event the iteration visits
        idx = np.arange(c - 1, first - 1, -1)              # newest -> oldest
        start = int(ts[lo]) if full else (newest - span_ns if newest > span_ns else 0)
        local = ts[idx].astype(np.int64) - start
        out.append((held, len(idx), newest, int(ts[first]), start, int(local.sum()),
                    int(((np.arange(len(idx)) + 1) * fr_x[idx].astype(np.int64)).sum())))
    return out


def test_golden_stream_slicing(sl):
    """Slice boundaries of the golden DVS_flow<50000, 200 ms> streams (minted from the compiled reference)."""
    G = json.load(open(os.path.join(HERE, "golden", "golden.json")))
    EV = np.load(os.path.join(HERE, "golden", "events.npz"))
    x, y, t = EV["stream_x"], EV["stream_y"], EV["stream_t_ns"].astype(np.uint64)
    for g in G["streams"]:
        info, rec = slices_of(sl, y, x, t, config=g["config"], ev_refresh=g["ev_refresh"], time_refresh_ns=g["time_refresh_ns"])
        assert len(info) == g["n_slices"]
        assert info.tolist() == g["info"]
        # a full buffer (50000) is iterated without its oldest element: 49999 events reach the optimiser
        assert [int(r[0]) for r in rec] == [i[1] - (1 if i[1] == 50000 else 0) for i in g["info"]]


@pytest.mark.parametrize("config,capacity,span_ns,rate,dur,ev_refresh,time_refresh_ns", [
    (0, 50000, 200_000_000, 1.0e6, 0.12, 20000, 33_000_000),     # CLI configuration: the buffer overflows (event trigger)
    (0, 50000, 200_000_000, 0.1e6, 0.50, 20000, 33_000_000),     # sparse stream: the time trigger fires, the span evicts
    (1, 30000, 70_000_000, 0.6e6, 0.20, 15000, 20_000_000),      # ROS node configuration: eviction by span and by size
])
def test_slicing_matches_the_compiled_reference(sl, config, capacity, span_ns, rate, dur, ev_refresh, time_refresh_ns):
    from oracle import ref
    if not ref.available(180, 240):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    st = synth.make_stream(240, 180, rate, dur, seed=41 + config)
    fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.uint64)
    _, want_info = ref.stream(fr_x, fr_y, ts, config=config, ev_refresh=ev_refresh, time_refresh_ns=time_refresh_ns, scale=3,
                              max_iter=1, stm_disable=True, flush=True)
    info, rec = slices_of(sl, fr_x, fr_y, ts, config=config, ev_refresh=ev_refresh, time_refresh_ns=time_refresh_ns)
    assert len(info) >= 3
    assert info.tolist() == want_info.tolist()            # events consumed, buffer size, buffer time diff per slice
    # ... and the slices' contents against the ring buffer restated in numpy
    exp = expected_slices(np.asarray(fr_x), np.asarray(ts).astype(np.int64), [int(i[0]) for i in info], capacity, span_ns)
    for i, r, e in zip(info, rec, exp):
        assert int(i[1]) == e[0]
        assert r.tolist() == list(e[1:])
    if config == 0 and rate >= 1e6:
        assert any(int(i[1]) == capacity for i in info)    # the overflow quirk was exercised


def test_run_time_sized_buffer_equals_template_sized(sl):
    """--max-events / --slice-time build the same ring buffer at run time."""
    st = synth.make_stream(240, 180, 0.8e6, 0.15, seed=5)
    a = slices_of(sl, st.y, st.x, st.t_ns, config=0)
    b = slices_of(sl, st.y, st.x, st.t_ns, config=2, capacity=50000, span_ns=200_000_000)
    assert a[0].tolist() == b[0].tolist() and a[1].tolist() == b[1].tolist()
    c = slices_of(sl, st.y, st.x, st.t_ns, config=2, capacity=10000, span_ns=30_000_000)
    assert max(int(i[1]) for i in c[0]) <= 10000 and max(int(i[2]) for i in c[0]) <= 30_000_000
