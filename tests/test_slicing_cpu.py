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


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def test_host_event_formulas_match_the_oracle(sl, oracle_port):
    """The per-event formulas the host mirror keeps (better_flow/event.h: set_local_time, project,
    project_4param_reinit, compute_uv -- used for set-up, -o output and OptimizerLocal's write-back) against the
    oracle restatement (itself pinned bit-for-bit on the compiled reference), to the last bit."""
    rng = np.random.default_rng(11)
    n = 5000
    fx = rng.integers(0, 180, n).astype(np.uint32)
    fy = rng.integers(0, 240, n).astype(np.uint32)
    t = rng.integers(-2_000_000, 200_000_000, n).astype(np.int64)       # local times may be negative (dvs_flow.h:187-190)
    pr_x = fx + rng.normal(0, 2.0, n)
    pr_y = fy + rng.normal(0, 2.0, n)
    args = (0.031, -0.052, 88.5, 121.25, 3.5e-4, -2.25e-3)
    want = oracle_port.project(fx, fy, t, pr_x, pr_y, *args)
    px, py = pr_x.copy(), pr_y.copy()
    nx, ny = np.zeros(n), np.zeros(n)
    sl.ev_project_4param(n, _p(fx, C.c_uint32), _p(fy, C.c_uint32), _p(t, C.c_longlong), _p(px, C.c_double), _p(py, C.c_double),
                         _p(nx, C.c_double), _p(ny, C.c_double), *[C.c_double(a) for a in args])
    for got, w in zip((px, py, nx, ny), want):
        assert np.array_equal(got, w)

    u, v = np.zeros(n), np.zeros(n)
    sl.ev_compute_uv(n, _p(nx, C.c_double), _p(ny, C.c_double), _p(u, C.c_double), _p(v, C.c_double))
    wu, wv = oracle_port.compute_uv(nx, ny)
    assert np.array_equal(u, wu) and np.array_equal(v, wv)

    # set_local_time + project(nx, ny): a pure translation is project_4param_reinit with no rotation / divergence
    ts = rng.integers(1_000_000_000, 1_200_000_000, n).astype(np.uint64)
    t0 = 1_050_000_000
    tl, qx, qy = np.zeros(n, dtype=np.int64), np.zeros(n), np.zeros(n)
    sl.ev_project(n, _p(fx, C.c_uint32), _p(fy, C.c_uint32), _p(ts, C.c_ulonglong), C.c_ulonglong(t0), C.c_double(0.04),
                  C.c_double(-0.09), _p(tl, C.c_longlong), _p(qx, C.c_double), _p(qy, C.c_double))
    assert np.array_equal(tl, ts.astype(np.int64) - t0) and tl.min() < 0 < tl.max()
    w = oracle_port.project(fx, fy, tl, fx.astype(np.float64), fy.astype(np.float64), 0.04, -0.09, 0.0, 0.0, 0.0, 0.0)
    assert np.array_equal(qx, w[0]) and np.array_equal(qy, w[1])
