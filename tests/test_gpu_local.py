"""GPU parity of the OptimizerLocal path (bf_batch_add_local / bf_local_minimize):
integer images and IEEE double control flow, so everything -- nx, ny, score, step count, non-zero pixel
count -- must equal the oracle / the golden records of the reference EXACTLY."""
import os
import subprocess

import numpy as np
import pytest

import better_flow_b200 as bf
from better_flow_b200 import synth
from test_oracle_local import local_case_events, local_golden

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "better_flow_b200", "bf_motion_compensator")


def check(got, want_state, want_rc, want_steps, want_nz):
    assert got["rc"] == want_rc
    names = ("nx", "ny", "score", "dnx", "dny", "dn_th")
    if want_rc == 0:
        assert [float(got[k]).hex() for k in names] == [float(w).hex() for w in want_state]
        if want_steps >= 0:
            assert got["steps"] == want_steps
        assert got["nz_cnt"] == want_nz
    else:
        assert got["nx"] == 0 and got["ny"] == 0


@pytest.mark.parametrize("name", [c["name"] for c in local_golden()["cases"]])
def test_local_equals_reference_golden(name):
    case = next(c for c in local_golden()["cases"] if c["name"] == name)
    fx, fy, t = local_case_events(case)
    ctx = bf.Context(case["rows"], case["cols"], 3, max_events=len(fx) + 64, max_slices=4, device=0)
    try:
        got = ctx.local_minimize(fx, fy, t, case["scale"])
        check(got, [float.fromhex(h) for h in case["state"]], case["rc"], case["steps"], case["image_nz"])
        if case["rc"] == 0:
            assert (got["img_rows"], got["img_cols"]) == (case["img_rows"], case["img_cols"])
    finally:
        ctx.close()


def test_local_equals_oracle_on_fresh_clouds_all_group_sizes(ctx240, oracle_port):
    for seed, vel, dur, scale in [(81, (30.0, 10.0), 0.02, 3), (82, (-90.0, 50.0), 0.04, 3), (83, (60.0, 60.0), 0.04, 1)]:
        st = synth.make_stream(240, 180, 1.5e6, dur, seed=seed, vel=vel)
        sl = synth.cut_slices(st, dur)[0]
        want = oracle_port.local_minimize(sl.fr_x, sl.fr_y, sl.t_ns, scale, want_image=True)
        for G in (0, 1, 3, 8):
            ctx240.set_option("group_size", G)
            got = ctx240.local_minimize(sl.fr_x, sl.fr_y, sl.t_ns, scale)
            check(got, [want[k] for k in ("nx", "ny", "score", "dnx", "dny", "dn_th")], want["rc"], want["steps"],
                  int((want["image"] > 0).sum()))
    ctx240.set_option("group_size", 0)


def test_mixed_batch_rolling_and_local_slices(ctx240, oracle_port):
    """Rolling and local slices share one persistent launch; each equals its single-slice result."""
    st = synth.make_stream(240, 180, 2.0e6, 0.12, seed=91)
    sls = synth.cut_slices(st, 0.02)
    singles = []
    for k, s in enumerate(sls):
        if k % 2:
            singles.append(("local", ctx240.local_minimize(s.fr_x, s.fr_y, s.t_ns, 3)))
        else:
            singles.append(("rolling", ctx240.minimize(s.fr_x, s.fr_y, s.t_ns, 3, 12)))
    ctx240.reset()
    for k, s in enumerate(sls):
        if k % 2:
            ctx240.add_local(s.fr_x, s.fr_y, s.t_ns, 3)
        else:
            ctx240.add(s.fr_x, s.fr_y, s.t_ns, 3, 12)
    ctx240.run()
    for k, (kind, want) in enumerate(singles):
        r = ctx240.result(k)
        if kind == "local":
            got = ctx240.local_view(r)
            assert got == want
            o = oracle_port.local_minimize(sls[k].fr_x, sls[k].fr_y, sls[k].t_ns, 3)
            assert (got["nx"], got["ny"], got["score"], got["steps"]) == (o["nx"], o["ny"], o["score"], o["steps"])
        else:
            assert r["iters"] == want["iters"] and r["model"].tobytes() == want["model"].tobytes()


def test_local_scale5_equals_oracle(ctx240, oracle_port):
    """OptimizerLocal at scale 5 (optimizer_sampler.cpp:124-149 with a 5 x 5 splat and cv::GaussianBlur(5 x 5)): the
    cell geometry with a 4-pixel halo.  Exact equality with the oracle (itself bit-equal to the compiled reference
    class at scale 5, tests/test_oracle_local.py), for several CTA groupings, plus a window that touches all four
    image borders (the reflect-101 fix-ups of the blur) and a mixed batch."""
    cases = [(74, (40.0, -20.0), 0.02), (75, (-70.0, 30.0), 0.03)]
    for seed, vel, dur in cases:
        st = synth.make_stream(240, 180, 1.5e6, dur, seed=seed, vel=vel)
        sl = synth.cut_slices(st, dur)[0]
        want = oracle_port.local_minimize(sl.fr_x, sl.fr_y, sl.t_ns, 5, want_image=True)
        for G in (0, 1, 5):
            ctx240.set_option("group_size", G)
            got = ctx240.local_minimize(sl.fr_x, sl.fr_y, sl.t_ns, 5)
            check(got, [want[k] for k in ("nx", "ny", "score", "dnx", "dny", "dn_th")], want["rc"], want["steps"],
                  int((want["image"] > 0).sum()))
    ctx240.set_option("group_size", 0)
    # a small dense window: every border row / column of the image carries events
    rng = np.random.default_rng(8)
    n = 30000
    fx = rng.integers(60, 101, n).astype(np.uint16); fy = rng.integers(90, 151, n).astype(np.uint16)
    t = np.sort(rng.integers(0, 20_000_000, n)).astype(np.int32)[::-1].copy()
    want = oracle_port.local_minimize(fx, fy, t, 5, want_image=True)
    got = ctx240.local_minimize(fx, fy, t, 5)
    check(got, [want[k] for k in ("nx", "ny", "score", "dnx", "dny", "dn_th")], want["rc"], want["steps"], int((want["image"] > 0).sum()))
    # rolling scale 5, local scale 5 and local scale 3 in one launch
    st = synth.make_stream(240, 180, 1.5e6, 0.06, seed=76)
    sls = synth.cut_slices(st, 0.02)[:3]
    ctx240.reset()
    ctx240.add(sls[0].fr_x, sls[0].fr_y, sls[0].t_ns, 5, 6)
    ctx240.add_local(sls[1].fr_x, sls[1].fr_y, sls[1].t_ns, 5)
    ctx240.add_local(sls[2].fr_x, sls[2].fr_y, sls[2].t_ns, 3)
    ctx240.run()
    for k, sc in ((1, 5), (2, 3)):
        o = oracle_port.local_minimize(sls[k].fr_x, sls[k].fr_y, sls[k].t_ns, sc)
        g = ctx240.local_view(ctx240.result(k))
        assert (g["nx"], g["ny"], g["score"], g["steps"]) == (o["nx"], o["ny"], o["score"], o["steps"])
    ex = oracle_port.minimize(sls[0].fr_x, sls[0].fr_y, sls[0].t_ns, scale=5, max_iter=6, accum_mode=1)
    assert ctx240.result(0)["iters"] == ex["iters"] and np.allclose(ctx240.result(0)["model"][7:11], ex["model"][7:11], rtol=1e-9, atol=0)


def test_cli_optimizer_local(tmp_path, oracle_port):
    """bf_motion_compensator --optimizer=local: DVS_flow drives the C++ OptimizerLocal mirror per slice."""
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "better_flow_b200"), "cli"])
    st = synth.make_stream(240, 180, 1.0e6, 0.12, seed=93, vel=(60.0, -30.0))
    rec = np.zeros(len(st), dtype=np.dtype([("t", "<u8"), ("x", "<u2"), ("y", "<u2"), ("p", "<u4")]))
    rec["t"], rec["x"], rec["y"], rec["p"] = st.t_ns, st.x, st.y, st.p
    binf = tmp_path / "s.bin"
    rec.tofile(binf)
    out = tmp_path / "flow.txt"
    r = subprocess.run([CLI, "--quiet", "--stm-disable", "--optimizer=local", "--refresh-time=0.04", "--refresh-event-count=1000000",
                        "--slice-time=0.04", "--max-events=100000", "--flow-out=%s" % out, str(binf)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    rows = np.loadtxt(out, ndmin=2)
    assert rows.shape[0] >= 3
    # first slice: triggered by the first event at t >= 40 ms; the ring holds the n0 newest events up to it
    # (older ones were evicted by the 40 ms span), newest first; local time = t - (now - span)
    n0 = int(rows[0, 1])
    i_t = int(np.searchsorted(st.t_ns, 40_000_000, side="left"))
    order = np.arange(i_t, i_t - n0, -1)
    now = int(st.t_ns[i_t])
    start = max(0, now - 40_000_000)
    want = oracle_port.local_minimize(st.y[order], st.x[order], (st.t_ns[order] - start).astype(np.int64), 3)
    assert rows[0, 4] == -want["nx"] and rows[0, 5] == -want["ny"] and int(rows[0, 2]) == want["steps"]
