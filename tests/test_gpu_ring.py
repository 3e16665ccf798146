"""bf_ring_* (include/bf_cuda.h): the device-resident slice ring of DVS_flow's default mode -- only new events are
uploaded, a slice is an index range of the ring, and a warm-start chain is stream-ordered device work (the model of
slice k is read from slice k-1's result record on the device).  Every slice must equal what bf_minimize gives for the
same events with the same init model handed over by the host (the round-1 path)."""
import numpy as np
import pytest

import better_flow_b200 as bf
from better_flow_b200 import synth
from helpers import golden, unhex, ring_slice, same_model

pytestmark = pytest.mark.gpu


def _host_chain(ctx, fr_x, fr_y, ts, consumed, max_iter, chain, capacity=50000, span=200_000_000):
    out, last = [], (np.zeros(11) if chain else None)
    for c in consumed:
        idx, start = ring_slice(ts, c, capacity, span)
        r = ctx.minimize(fr_x[idx], fr_y[idx], (ts[idx] - start).astype(np.int32), 3, max_iter, init=last)
        if chain:
            last = r["model"]
        out.append(r)
    return out


def _ring_chain(ctx, fr_x, fr_y, ts, consumed, max_iter, chain, capacity=50000, span=200_000_000, max_pending=8, in_place=False):
    ring = bf.Ring(ctx, capacity, max_pending)
    try:
        tickets, fed = [], 0
        for c in consumed:
            if in_place:                                            # written straight into the pinned staging buffer, in two
                mid = fed + (c - fed) // 3                          # commits, each smaller than its reservation
                ring.push_in_place(fr_x[fed:mid], fr_y[fed:mid], ts[fed:mid], reserve=32768)
                ring.push_in_place(fr_x[mid:c], fr_y[mid:c], ts[mid:c], reserve=32768)
            else:
                ring.push(fr_x[fed:c], fr_y[fed:c], ts[fed:c])      # only the NEW events
            fed = c
            idx, start = ring_slice(ts, c, capacity, span)
            tickets.append(ring.slice(len(idx), start, 3, max_iter, chain))   # returns at once: nothing waits for the previous slice
        assert ring.pushed == fed
        return [ring.result(t) for t in tickets[-max_pending:]], len(tickets) - min(len(tickets), max_pending)
    finally:
        ring.close()


@pytest.mark.parametrize("ring_cluster", [0, 8, 16])
@pytest.mark.parametrize("chain", [True, False])
def test_ring_chain_equals_host_driven_chain(ctx240, chain, ring_cluster):
    st = synth.make_stream(240, 180, 1.5e6, 0.2, seed=77, vel=(60.0, 35.0), omega=0.4)
    fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.int64)
    consumed = list(range(20000, len(ts), 20000))
    ctx240.set_option("ring_cluster", 0)
    want = _host_chain(ctx240, fr_x, fr_y, ts, consumed, -1, chain)
    ctx240.set_option("ring_cluster", ring_cluster)       # the single-slice launches of the ring on one thread-block cluster
    try:
        got, first = _ring_chain(ctx240, fr_x, fr_y, ts, consumed, -1, chain, max_pending=len(consumed))
    finally:
        ctx240.set_option("ring_cluster", 0)
    assert first == 0 and len(got) == len(want) >= 10
    for w, g in zip(want, got):
        assert g["rc"] == w["rc"] == 0 and g["iters"] == w["iters"] and g["n_events"] == w["n_events"]
        assert same_model(g["model"], w["model"])
        assert g["dividers"].tobytes() == w["dividers"].tobytes()


def test_ring_reserve_commit_equals_push(ctx240):
    """bf_ring_reserve / bf_ring_commit (what DVS_flow::add_event writes through): same ring content, same chain; a
    small ring (capacity 30000) so that staging and device ring both wrap; an event outside the sensor is refused at
    commit and leaves the ring usable."""
    st = synth.make_stream(240, 180, 1.5e6, 0.3, seed=79, vel=(40.0, -55.0), omega=0.2)
    fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.int64)
    consumed = list(range(20000, len(ts), 20000))
    kw = dict(capacity=30000, span=70_000_000, max_pending=len(consumed))
    want, _ = _ring_chain(ctx240, fr_x, fr_y, ts, consumed, 8, True, **kw)
    got, _ = _ring_chain(ctx240, fr_x, fr_y, ts, consumed, 8, True, in_place=True, **kw)
    assert len(got) == len(want) >= 15
    for w, g in zip(want, got):
        assert g["iters"] == w["iters"] and g["n_events"] == w["n_events"] and same_model(g["model"], w["model"])
    # one commit larger than the ring: only its newest `capacity` events can ever be used
    res = []
    for in_place in (False, True):
        ring = bf.Ring(ctx240, 12000, 4)
        try:
            (ring.push_in_place if in_place else ring.push)(fr_x[:30000], fr_y[:30000], ts[:30000])
            assert ring.pushed == 30000
            res.append(ring.result(ring.slice(12000, int(ts[18000]), 3, 6, False)))
        finally:
            ring.close()
    w = ctx240.minimize(fr_x[18000:30000][::-1], fr_y[18000:30000][::-1], (ts[18000:30000][::-1] - ts[18000]).astype(np.int32), 3, 6)
    for g in res:
        assert g["iters"] == w["iters"] and g["n_events"] == 12000 and same_model(g["model"], w["model"])
    ring = bf.Ring(ctx240, 30000, 4)
    try:
        with pytest.raises(bf.BfError):
            ring.push_in_place(np.array([10, 500]), np.array([10, 10]), np.array([1, 2]))
        assert ring.pushed == 0
        ring.push_in_place(fr_x[:5000], fr_y[:5000], ts[:5000])
        assert ring.pushed == 5000
        with pytest.raises(bf.BfError):
            ring._chk(ring.lib.bf_ring_commit(ring.h, 1))                  # no reservation open
    finally:
        ring.close()


def test_ring_seed_continues_a_chain_in_a_new_ring(ctx240):
    """bf_ring_seed: a ring created in the middle of a stream (DVS_flow after a context was re-created or after the host
    path handled some slices) starts its chain from the caller's last model, not from ObjectModel()."""
    st = synth.make_stream(240, 180, 1.5e6, 0.12, seed=78, vel=(-50.0, 45.0), omega=-0.3)
    fr_x, fr_y, ts = st.y, st.x, st.t_ns.astype(np.int64)
    consumed = list(range(20000, len(ts), 20000))
    want = _host_chain(ctx240, fr_x, fr_y, ts, consumed, -1, True)
    k = len(consumed) // 2
    assert k >= 2
    ring = bf.Ring(ctx240, 50000, 8)
    try:
        lo = max(0, consumed[k] - 50000)
        ring.push(fr_x[lo:consumed[k]], fr_y[lo:consumed[k]], ts[lo:consumed[k]])   # the window as slice k sees it
        ring.seed(want[k - 1]["model"])
        got, fed = [], consumed[k]
        for c in consumed[k:]:
            ring.push(fr_x[fed:c], fr_y[fed:c], ts[fed:c])
            fed = c
            idx, start = ring_slice(ts, c)
            got.append(ring.result(ring.slice(len(idx), start, 3, -1, True)))
    finally:
        ring.close()
    for w, g in zip(want[k:], got):
        assert g["iters"] == w["iters"] and same_model(g["model"], w["model"])
    # (without the seed the first of these slices starts cold and takes a different number of steps)
    ring = bf.Ring(ctx240, 50000, 8)
    try:
        ring.push(fr_x[lo:consumed[k]], fr_y[lo:consumed[k]], ts[lo:consumed[k]])
        idx, start = ring_slice(ts, consumed[k])
        cold = ring.result(ring.slice(len(idx), start, 3, -1, True))
    finally:
        ring.close()
    assert not same_model(cold["model"], want[k]["model"])


def test_ring_golden_warm_start_stream(ctx240):
    """The golden DVS_flow<50000, 200 ms> warm-start stream (minted from the compiled reference): the device chain
    within the free-running tolerance of tests/test_gpu_cli.py, the first slice within 1e-6."""
    G, EV = golden()
    g = [s for s in G["streams"] if not s["stm_disable"]][0]
    fr_x, fr_y, ts = EV["stream_y"], EV["stream_x"], EV["stream_t_ns"].astype(np.int64)
    got, _ = _ring_chain(ctx240, fr_x, fr_y, ts, [i[0] for i in g["info"]], g["max_iter"], True)
    want = [unhex(m) for m in g["models"]]
    rel = np.array([np.max(np.abs(r["model"][7:9] - m[7:9]) / np.abs(m[7:9])) for r, m in zip(got, want)])
    assert rel[0] < 1e-6 and np.all(rel < 6e-4), rel


def test_ring_tiny_window_marks_noise_on_the_device(ctx240):
    """A slice that trips the tiny-window guard marks its events as noise (optimizer_rolling.h:49-55); overlapping
    later slices must skip them -- here that happens on the device, in the ring."""
    rng = np.random.default_rng(3)
    n0 = 12000
    fx0, fy0 = rng.integers(80, 88, n0), rng.integers(100, 110, n0)
    st = synth.make_stream(240, 180, 1.0e6, 0.06, seed=14)
    fr_x = np.concatenate([fx0, st.y]).astype(np.uint16); fr_y = np.concatenate([fy0, st.x]).astype(np.uint16)
    ts = np.concatenate([np.sort(rng.integers(0, 10 ** 7, n0)), st.t_ns + 10 ** 7]).astype(np.int64)
    consumed = [10000, 20000, 30000, 40000, 50000]
    got, _ = _ring_chain(ctx240, fr_x, fr_y, ts, consumed, 4, False)
    # host-driven reference of the same thing: the noise flags carried by hand
    noise = np.zeros(len(ts), dtype=np.uint8)
    for c, g in zip(consumed, got):
        idx, start = ring_slice(ts, c)
        w = ctx240.minimize(fr_x[idx], fr_y[idx], (ts[idx] - start).astype(np.int32), 3, 4, noise=noise[idx])
        if w["flags"] & bf.FLAG_ALL_NOISE:
            noise[idx] = 1
        assert g["rc"] == w["rc"] and g["iters"] == w["iters"] and g["flags"] == w["flags"]
        assert same_model(g["model"], w["model"]) or (g["rc"] == bf.RC_SKIPPED and np.array_equal(g["model"], w["model"]))
    assert got[0]["flags"] & bf.FLAG_ALL_NOISE and got[-1]["rc"] == 0


def test_ring_argument_checks(ctx240):
    ring = bf.Ring(ctx240, 1000, 4)
    try:
        with pytest.raises(bf.BfError):
            ring.slice(10, 0)                                       # nothing pushed yet
        ring.push(np.arange(10) % 180, np.arange(10) % 240, np.arange(10) * 1000)
        with pytest.raises(bf.BfError):
            ring.push(np.array([500]), np.array([1]), np.array([1]))   # outside the sensor
        t = ring.slice(10, 0, 3, 2, False)
        assert ring.result(t)["rc"] == bf.RC_SKIPPED and ring.result(t)["n_events"] == 10
        with pytest.raises(bf.BfError):
            ring.result(t + 1)
    finally:
        ring.close()
