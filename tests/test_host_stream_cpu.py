"""The HOST side of the drop-in against the reference's own DVS_flow, without a GPU: the mirror classes
(better_flow_b200/include/better_flow: DVS_flow, OptimizerRolling, OptimizerLocal, Event, CircularArray) are linked
against a test double of the C ABI whose compute is the oracle (tests/cpu/mock_bf_cuda.cpp; the oracle's rolling path
is pinned bit for bit on the compiled reference), so whatever differs from the reference's per-slice models is a bug in
the host plumbing: slice hand-over order, local times, warm start through last_model (centre used as stored), noise
marking, batching of independent slices."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from better_flow_b200 import synth
from helpers import unhex

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hs():
    so = os.path.join(HERE, "cpu", "libstream_shim.so")
    srcs = [os.path.join(HERE, "cpu", f) for f in ("stream_shim.cpp", "mock_bf_cuda.cpp")]
    inc = os.path.join(ROOT, "better_flow_b200", "include")
    odir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["make", "-s", "-C", odir, "port"])
    deps = srcs + [os.path.join(inc, "better_flow", f) for f in os.listdir(os.path.join(inc, "better_flow"))] + \
        [os.path.join(ROOT, "include", "bf_cuda.h"), os.path.join(odir, "libbf_oracle.so")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in deps):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I" + inc,
                               "-I" + os.path.join(ROOT, "include"), *srcs, "-L" + odir, "-lbf_oracle", "-Wl,-rpath," + odir, "-o", so])
    lib = C.CDLL(so)
    lib.bf_mock_minimize_calls.restype = C.c_longlong
    lib.bf_mock_batch_runs.restype = C.c_longlong
    lib.bf_mock_ring_pushed_events.restype = C.c_longlong
    lib.bf_mock_ring_slices.restype = C.c_longlong
    return lib


def run_host(lib, fr_x, fr_y, ts, config=0, ev_refresh=20000, time_refresh_ns=33000000, scale=3, max_iter=-1, stm_disable=False,
             flush=True, batch=1, local=False, lazy=False, max_slices=512):
    fx = np.ascontiguousarray(fr_x, dtype=np.uint32)
    fy = np.ascontiguousarray(fr_y, dtype=np.uint32)
    t = np.ascontiguousarray(ts, dtype=np.uint64)
    models = np.zeros((max_slices, 11)); info = np.zeros((max_slices, 3), dtype=np.int64); uv = np.zeros((max_slices, 14))
    p = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
    k = lib.st_stream(config, len(fx), p(fx, C.c_uint32), p(fy, C.c_uint32), p(t, C.c_uint64), C.c_ulonglong(ev_refresh),
                      C.c_ulonglong(time_refresh_ns), scale, max_iter, 1 if stm_disable else 0, 1 if flush else 0, batch,
                      1 if local else 0, int(lazy), max_slices, p(models, C.c_double), p(info, C.c_longlong), p(uv, C.c_double))
    assert 0 <= k <= max_slices
    return models[:k], info[:k], uv[:k]


def test_golden_streams_bit_for_bit(hs):
    """tests/golden: DVS_flow<50000, 200 ms> of the compiled reference, warm-start chain and --stm-disable."""
    G = json.load(open(os.path.join(HERE, "golden", "golden.json")))
    EV = np.load(os.path.join(HERE, "golden", "events.npz"))
    x, y, t = EV["stream_x"], EV["stream_y"], EV["stream_t_ns"].astype(np.uint64)
    for g in G["streams"]:
        models, info, _ = run_host(hs, y, x, t, config=g["config"], ev_refresh=g["ev_refresh"], time_refresh_ns=g["time_refresh_ns"],
                                   scale=g["scale"], max_iter=g["max_iter"], stm_disable=g["stm_disable"])
        want = np.stack([unhex(m) for m in g["models"]])
        assert info.tolist() == g["info"]
        assert models.shape == want.shape
        assert np.array_equal(models, want), (g["stm_disable"], np.abs(models - want).max())


@pytest.mark.parametrize("config,rate,dur,stm_disable,scale,max_iter", [
    (0, 1.0e6, 0.10, False, 3, 6),        # the tool's configuration, warm-start chain, buffer overflow
    (0, 1.0e6, 0.10, True, 3, 6),
    (1, 0.6e6, 0.15, False, 3, 4),        # the ROS node's configuration
    (0, 0.8e6, 0.08, False, 5, 3),        # scale 5
    (0, 0.8e6, 0.08, True, 1, 8),         # scale 1
])
def test_host_mirror_equals_the_compiled_reference(hs, config, rate, dur, stm_disable, scale, max_iter):
    from oracle import ref
    if not ref.available(180, 240):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    st = synth.make_stream(240, 180, rate, dur, seed=60 + config + scale, vel=(70.0, -110.0), omega=0.6, expand=0.3)
    fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.uint64)
    want_models, want_info = ref.stream(fr_x, fr_y, ts, config=config, ev_refresh=15000, time_refresh_ns=25_000_000, scale=scale,
                                        max_iter=max_iter, stm_disable=stm_disable, flush=True)
    models, info, uv = run_host(hs, fr_x, fr_y, ts, config=config, ev_refresh=15000, time_refresh_ns=25_000_000, scale=scale,
                                max_iter=max_iter, stm_disable=stm_disable)
    assert len(models) == len(want_models) >= 3
    assert info.tolist() == want_info.tolist()
    assert np.array_equal(models, want_models), np.abs(models - want_models).max()
    assert np.all(np.isfinite(uv)) and np.any(uv[:, :2] != 0)       # compute_uv ran over the buffer after every slice


def test_batched_independent_slices_equal_unbatched(hs):
    """--stm-disable --batch=N queues slices and minimises them together: same models, in order, fewer back-end calls."""
    st = synth.make_stream(240, 180, 1.0e6, 0.12, seed=91)
    fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.uint64)
    c0, b0 = hs.bf_mock_minimize_calls(), hs.bf_mock_batch_runs()
    a, info, _ = run_host(hs, fr_x, fr_y, ts, stm_disable=True, max_iter=5)
    c1 = hs.bf_mock_minimize_calls()
    b, binfo, _ = run_host(hs, fr_x, fr_y, ts, stm_disable=True, max_iter=5, batch=3)
    assert len(a) == len(b) >= 5 and c1 - c0 == len(a)
    assert hs.bf_mock_minimize_calls() == c1                         # batched: no per-slice calls ...
    assert hs.bf_mock_batch_runs() - b0 == -(-len(a) // 3)           # ... but ceil(n / 3) batch runs
    assert np.array_equal(a, b)
    assert binfo[:, 1].tolist() == [min(int(i), 49999) if int(i) == 50000 else int(i) for i in info[:, 1]]


def test_lazy_events_change_no_model(hs):
    """set_lazy_events (the CLI without -o): the per-event state is not read back, the per-slice models -- warm-start
    chain included -- are the same; the tiny-window guard still marks its events as noise for the slices that follow."""
    st = synth.make_stream(240, 180, 1.0e6, 0.12, seed=33, vel=(-60.0, 90.0))
    fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.uint64)
    for stm in (False, True):
        a, ia, uva = run_host(hs, fr_x, fr_y, ts, stm_disable=stm, max_iter=6)
        b, ib, uvb = run_host(hs, fr_x, fr_y, ts, stm_disable=stm, max_iter=6, lazy=True)
        assert np.array_equal(a, b) and ia.tolist() == ib.tolist()
        assert np.any(uva[:, 0] != 0) and np.all(uvb[:, 0] == 0) and np.all(uvb[:, 4] == 0)   # u, nx stay as Event::reset left them
        assert np.array_equal(uva[:, 11:], uvb[:, 11:])                                       # local times, noise flags, counts
    rng = np.random.default_rng(1)
    n = 40000
    fx, fy = rng.integers(80, 88, n), rng.integers(100, 110, n)          # everything inside a tiny window: all-noise guard
    t = np.sort(rng.integers(10 ** 9, 10 ** 9 + 10 ** 8, n)).astype(np.uint64)
    a, _, uva = run_host(hs, fx, fy, t, ev_refresh=10000, max_iter=3)
    b, _, uvb = run_host(hs, fx, fy, t, ev_refresh=10000, max_iter=3, lazy=True)
    assert np.array_equal(a, b) and np.array_equal(uva[:, 12], uvb[:, 12]) and uva[1, 12] > 0


def test_optimizer_local_mode_against_the_compiled_class(hs):
    """--optimizer=local: every slice goes through the OptimizerLocal mirror (cloud copy, run, write-back); the slice
    model carries -nx, -ny and the score.  Checked against the compiled reference class run on the same slices
    (reconstructed here from the ring-buffer rules that test_slicing_cpu.py pins)."""
    from oracle import ref
    if not ref.available(180, 240):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    st = synth.make_stream(240, 180, 0.5e6, 0.09, seed=71, vel=(120.0, 40.0))
    fr_x, fr_y, ts = np.asarray(st.y), np.asarray(st.x), np.asarray(st.t_ns).astype(np.int64)
    models, info, _ = run_host(hs, fr_x, fr_y, ts.astype(np.uint64), ev_refresh=9000, time_refresh_ns=33_000_000, scale=3, local=True)
    assert len(models) >= 4
    cap, span = 50000, 200_000_000
    for m, i in zip(models, info):
        c = int(i[0])
        newest = int(ts[c - 1])
        lo = int(np.searchsorted(ts[:c], newest - span, side="right")) if newest >= span else 0
        lo = max(lo, c - cap)
        full = (c - lo) == cap
        first = lo + (1 if full else 0)
        start = int(ts[lo]) if full else (newest - span if newest > span else 0)
        idx = np.arange(c - 1, first - 1, -1)
        want = ref.local_minimize(fr_x[idx], fr_y[idx], ts[idx] - start, scale=3)
        assert m[7] == -want["nx"] and m[8] == -want["ny"] and m[2] == want["score"], (i.tolist(), m[7:9], want["nx"], want["ny"])
    assert np.any(models[:, 7] != 0)


def test_batched_tiny_window_marks_noise_like_the_unbatched_path(hs):
    """--stm-disable --batch=N: a slice that trips the tiny-window guard (optimizer_rolling.h:49-55) must mark its
    events as noise in the live buffer BEFORE the next overlapping slice is snapshotted, exactly as the unbatched
    path and the reference do -- the batch itself is minimised later."""
    rng = np.random.default_rng(3)
    n0 = 12000
    fx0, fy0 = rng.integers(80, 88, n0), rng.integers(100, 110, n0)     # first 12 k events: a tiny window (-> all noise)
    st = synth.make_stream(240, 180, 1.0e6, 0.06, seed=14)              # then a normal scene, overlapping slices
    fx = np.concatenate([fx0, st.y]); fy = np.concatenate([fy0, st.x])
    t = np.concatenate([np.sort(rng.integers(10 ** 9, 10 ** 9 + 10 ** 7, n0)), st.t_ns + 10 ** 9 + 10 ** 7]).astype(np.uint64)
    a, ia, _ = run_host(hs, fx, fy, t, ev_refresh=10000, stm_disable=True, max_iter=4)
    b, ib, _ = run_host(hs, fx, fy, t, ev_refresh=10000, stm_disable=True, max_iter=4, batch=4)
    assert len(a) == len(b) >= 5
    assert np.array_equal(a, b), np.abs(a - b).max()
    assert np.all(a[1][7:] == 0) and np.any(a[-1] != 0)                  # the guard fired early on, later slices ran
    from oracle import ref
    if ref.available(180, 240):
        want, _ = ref.stream(fx, fy, t, config=0, ev_refresh=10000, scale=3, max_iter=4, stm_disable=True, flush=True)
        assert np.array_equal(b, want)


@pytest.mark.parametrize("config,stm", [(0, False), (0, True), (1, False)])
def test_device_ring_mode_changes_no_model(hs, config, stm):
    """set_device_ring (the CLI's default without -o): the ring lives behind the C ABI, the host pushes only the NEW
    events of every slice and enqueues "the newest n events" -- warm starts chain behind the ABI, models are read back
    late.  Same models as the host-driven path; exactly the stream's events cross the ABI once."""
    st = synth.make_stream(240, 180, 1.0e6, 0.16, seed=47, vel=(-60.0, 90.0), omega=0.5)
    fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.uint64)
    a, ia, _ = run_host(hs, fr_x, fr_y, ts, config=config, stm_disable=stm, max_iter=6, lazy=1)
    c0, p0, s0 = hs.bf_mock_minimize_calls(), hs.bf_mock_ring_pushed_events(), hs.bf_mock_ring_slices()
    b, ib, uvb = run_host(hs, fr_x, fr_y, ts, config=config, stm_disable=stm, max_iter=6, lazy=2)
    assert len(a) == len(b) >= 6 and ia.tolist() == ib.tolist()
    assert np.array_equal(a, b), np.abs(a - b).max()
    assert hs.bf_mock_minimize_calls() == c0                               # no per-slice hand-over any more
    assert hs.bf_mock_ring_slices() - s0 == len(b)
    assert hs.bf_mock_ring_pushed_events() - p0 == len(ts)                 # every event uploaded exactly once
    from oracle import ref
    if ref.available(180, 240):
        want, _ = ref.stream(fr_x, fr_y, ts, config=config, scale=3, max_iter=6, stm_disable=stm, flush=True)
        assert np.array_equal(b, want)


@pytest.mark.parametrize("stm", [False, True])
def test_device_ring_switched_off_and_on_mid_stream(hs, stm):
    """The window lives in ONE of two host rings (152-byte Events, or the device ring's 16-byte records while the slices
    are cut behind the ABI).  Switching the mode with events in the window moves them across (size, order and the
    full-buffer quirk preserved), reads the outstanding models back first and rebuilds the device ring when the mode
    returns: no slice, trigger or model changes."""
    st = synth.make_stream(240, 180, 1.0e6, 0.2, seed=52, vel=(70.0, -40.0), omega=-0.4)
    fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.uint64)
    a, ia, _ = run_host(hs, fr_x, fr_y, ts, stm_disable=stm, max_iter=5, lazy=1)
    p0, s0 = hs.bf_mock_ring_pushed_events(), hs.bf_mock_ring_slices()
    b, ib, _ = run_host(hs, fr_x, fr_y, ts, stm_disable=stm, max_iter=5, lazy=3)
    assert len(a) == len(b) >= 8 and ia.tolist() == ib.tolist()
    assert np.array_equal(a, b), np.abs(a - b).max()
    assert 0 < hs.bf_mock_ring_slices() - s0 < len(b)                      # some slices on each side of the switches
    assert hs.bf_mock_ring_pushed_events() - p0 > 0


@pytest.mark.parametrize("stm", [False, True])
def test_device_ring_survives_a_recreated_context(hs, stm):
    """Another user of the pooled context asks for more capacity in the middle of the stream: the context is destroyed
    and re-created, and the ring and the pinned staging buffer add_event was writing into go with it.  The events that
    arrive before the next slice must not be written into the freed buffer (per-event generation check); the next slice
    rebuilds the ring from the host's window and seeds the chain with the host's last model: no model changes."""
    st = synth.make_stream(240, 180, 1.0e6, 0.2, seed=53, vel=(-45.0, 65.0), omega=0.2)
    fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.uint64)
    a, ia, _ = run_host(hs, fr_x, fr_y, ts, stm_disable=stm, max_iter=5, lazy=1)
    b, ib, _ = run_host(hs, fr_x, fr_y, ts, stm_disable=stm, max_iter=5, lazy=4)
    assert len(a) == len(b) >= 8 and ia.tolist() == ib.tolist()
    assert np.array_equal(a, b), np.abs(a - b).max()


def test_device_ring_more_new_events_than_one_staging_reservation(hs):
    """New events go straight into the ring's pinned staging buffer, 32768 at a reservation (bf_ring_reserve / commit).
    With 45000 new events per slice (only the time trigger of 100 ms could fire earlier) a reservation fills between
    two slices and is committed and renewed from add_event itself; with the window capacity of 30000 (config 1) a slice
    even brings more new events than the ring can hold.  Same models, no event uploaded twice."""
    st = synth.make_stream(240, 180, 1.2e6, 0.2, seed=61, vel=(30.0, 60.0))
    fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.uint64)
    for config in (0, 1):
        kw = dict(config=config, ev_refresh=45000, time_refresh_ns=100_000_000, max_iter=4)
        a, ia, _ = run_host(hs, fr_x, fr_y, ts, lazy=1, **kw)
        p0 = hs.bf_mock_ring_pushed_events()
        b, ib, _ = run_host(hs, fr_x, fr_y, ts, lazy=2, **kw)
        assert len(a) == len(b) >= 4 and ia.tolist() == ib.tolist()
        assert np.array_equal(a, b)
        # (the first slice builds the ring from the window: with config 1 that is the newest 30000 of its 45000 events)
        assert hs.bf_mock_ring_pushed_events() - p0 == len(ts) - (15000 if config == 1 else 0)


def test_device_ring_tiny_window_noise(hs):
    """The tiny-window guard's noise marks live behind the ABI in ring mode: later overlapping slices skip the events."""
    rng = np.random.default_rng(3)
    n0 = 12000
    fx0, fy0 = rng.integers(80, 88, n0), rng.integers(100, 110, n0)
    st = synth.make_stream(240, 180, 1.0e6, 0.06, seed=14)
    fx = np.concatenate([fx0, st.y]); fy = np.concatenate([fy0, st.x])
    t = np.concatenate([np.sort(rng.integers(10 ** 9, 10 ** 9 + 10 ** 7, n0)), st.t_ns + 10 ** 9 + 10 ** 7]).astype(np.uint64)
    for stm in (False, True):
        a, _, _ = run_host(hs, fx, fy, t, ev_refresh=10000, stm_disable=stm, max_iter=4, lazy=1)
        b, _, _ = run_host(hs, fx, fy, t, ev_refresh=10000, stm_disable=stm, max_iter=4, lazy=2)
        assert len(a) == len(b) >= 5 and np.array_equal(a, b)
