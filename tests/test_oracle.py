"""CPU tests that PIN the oracle (oracle/bf_oracle.c, the C restatement of the hot path):

  * against the golden vectors in tests/golden/, which were minted from the reference's own
    unmodified sources (oracle/make_golden.py) -- bit for bit: model scalars as hex floats,
    iteration counts, dividers, SHA-256 of every per-event output and of the stage-level images;
  * against the compiled reference itself (oracle/_ref) when that library is present, on fresh
    random slices -- also bit for bit;
  * and the exact-sum accumulation mode (what the CUDA path computes) against the reference mode
    within the tolerance ladder of SURVEY.md 8(c).
"""
import numpy as np
import pytest

from better_flow_b200 import synth
from helpers import case_events, golden, rel, sha, stage_positions, unhex

G, EV = golden()


def test_fixture_integrity():
    for k, h in G["event_sha"].items():
        assert sha(EV[k]) == h, k


@pytest.mark.parametrize("case", G["cases"], ids=[c["name"] for c in G["cases"]])
def test_port_matches_golden_bit_for_bit(oracle_port, case):
    fx, fy, t, noise, init = case_events(case)
    r = oracle_port.minimize(fx, fy, t, scale=case["scale"], max_iter=case["max_iter"], init_model=init, noise=noise,
                             rows=case["rows"], cols=case["cols"], accum_mode=0, want_events=True)
    assert r["rc"] == case["rc"]
    assert r["iters"] == case["iters"]
    assert [float(v).hex() for v in r["model"]] == case["model"]
    assert [float(v) for v in r["dividers"]] == case["dividers"]
    for k, v in case["setup"].items():
        assert r[k] == v, k
    for k in ("pr_x", "pr_y", "nx", "ny"):
        assert sha(r[k]) == case["sha_" + k], k
    assert int(r["noise"].sum()) == case["noise_out_sum"]


@pytest.mark.parametrize("st", G["stages"], ids=["scale%d" % s["scale"] for s in G["stages"]])
def test_port_stage_level_matches_golden(oracle_port, st):
    fx, fy, t = EV["davis240_a_fr_x"], EV["davis240_a_fr_y"], EV["davis240_a_t_ns"]
    pr_x, pr_y = stage_positions(fx, fy)
    img = oracle_port.time_img(pr_x, pr_y, t, st["w"], st["h"], st["scale"], st["x_sh"], st["y_sh"], accum_mode=0)
    assert sha(img) == st["sha_img"]
    assert int((img > 0).sum()) == st["nnz"]
    m7, gx, gy = oracle_port.model(img, want_grad=True)
    assert [float(v).hex() for v in m7] == st["model7"]
    assert sha(gx) == st["sha_gx"] and sha(gy) == st["sha_gy"]
    # the exact-sum mode differs from the reference accumulation only by f32 rounding noise
    exact = oracle_port.time_img(pr_x, pr_y, t, st["w"], st["h"], st["scale"], st["x_sh"], st["y_sh"], accum_mode=1)
    assert np.array_equal(exact > 0, img > 0)
    assert np.max(np.abs(exact - img)) < 1e-6


def test_port_projection_matches_golden(oracle_port):
    fx, fy, t = EV["davis240_a_fr_x"], EV["davis240_a_fr_y"], EV["davis240_a_t_ns"]
    pr_x, pr_y = stage_positions(fx, fy)
    args = unhex(G["project"]["args"])
    px, py, nx, ny = oracle_port.project(fx, fy, t, pr_x, pr_y, *args)
    for name, a in (("pr_x", px), ("pr_y", py), ("nx", nx), ("ny", ny)):
        assert sha(a) == G["project"]["sha_" + name], name


def test_exact_sum_mode_within_tolerance_ladder(oracle_port):
    """final (dx,dy) must stay far inside the 1e-4 contract when only the accumulation
    order / rounding of the time image changes."""
    for case in G["cases"]:
        if case["rc"] != 0:
            continue
        fx, fy, t, noise, init = case_events(case)
        r = oracle_port.minimize(fx, fy, t, scale=case["scale"], max_iter=case["max_iter"], init_model=init, noise=noise,
                                 rows=case["rows"], cols=case["cols"], accum_mode=1)
        want = unhex(case["model"])
        assert r["iters"] == case["iters"], case["name"]
        assert np.all(rel(r["model"][7:9], want[7:9]) < 1e-6), (case["name"], rel(r["model"][7:9], want[7:9]))


def test_port_equals_compiled_reference_on_fresh_slices(oracle_port):
    from oracle import ref
    if not ref.available(180, 240):
        pytest.skip("oracle/_ref not built (needs /root/reference); golden vectors pin the oracle instead")
    st = synth.make_stream(240, 180, 2.5e6, 0.024, seed=77, vel=(55.0, 95.0), omega=-0.7, expand=-0.3)
    for sl, (scale, mi) in zip(synth.cut_slices(st, 0.012)[:2], [(3, 14), (1, -1)]):
        a = ref.minimize(sl.fr_x, sl.fr_y, sl.t_ns, scale=scale, max_iter=mi, want_events=True)
        b = oracle_port.minimize(sl.fr_x, sl.fr_y, sl.t_ns, scale=scale, max_iter=mi, want_events=True)
        assert a["iters"] == b["iters"] and a["rc"] == b["rc"]
        assert np.array_equal(a["model"], b["model"])
        assert np.array_equal(a["dividers"], b["dividers"])
        for k in ("pr_x", "pr_y", "nx", "ny"):
            assert np.array_equal(a[k], b[k]), k


def test_compute_uv(oracle_port):
    # event.h:135-142 with the integer constant 1000000000/(T_DIVIDER*10000) = 100000
    nx = np.array([0.0, 0.127, -0.254, 0.05])
    ny = np.array([0.0, 0.0, 0.127, -0.07])
    u, v = oracle_port.compute_uv(nx, ny)
    assert u[0] == 0 and v[0] == 0
    assert np.allclose(u, nx * 100000 / 127.0, rtol=1e-12)
    assert np.allclose(v, ny * 100000 / 127.0, rtol=1e-12)


def test_port_equals_compiled_reference_on_randomised_slices(oracle_port):
    """A seeded sweep of slice shapes the goldens do not hold: all three scales, capped and converged runs, noise flags,
    warm starts, negative local times, windows that cover part of the sensor, slices below the 1000-event guard, the
    DAVIS-346 build -- the C restatement must equal the compiled reference to the last bit on every one."""
    from oracle import ref
    if not (ref.available(180, 240) and ref.available(260, 346)):
        pytest.skip("oracle/_ref not built (needs /root/reference); golden vectors pin the oracle instead")
    rng = np.random.default_rng(8)
    seen_rc = set()
    for k in range(24):
        rows, cols = (260, 346) if k % 6 == 5 else (180, 240)
        st = synth.make_stream(cols, rows, float(rng.uniform(0.4e6, 1.2e6)), 0.01, seed=200 + k,
                               vel=(float(rng.uniform(-150, 150)), float(rng.uniform(-150, 150))),
                               omega=float(rng.uniform(-1.5, 1.5)) if k % 3 == 0 else 0.0,
                               expand=float(rng.uniform(-0.8, 0.8)) if k % 4 == 0 else 0.0)
        sl = synth.cut_slices(st, 0.01)[0]
        fx, fy, t = np.asarray(sl.fr_x), np.asarray(sl.fr_y), np.asarray(sl.t_ns).astype(np.int64)
        if k % 5 == 1:                                            # a window in the middle of the sensor
            keep = (fx > rows // 4) & (fx < 3 * rows // 4) & (fy > cols // 5) & (fy < 4 * cols // 5)
            fx, fy, t = fx[keep], fy[keep], t[keep]
        if k % 7 == 3:                                            # below the 1000-event guard (optimizer_rolling.h:57-58)
            fx, fy, t = fx[:700], fy[:700], t[:700]
        if k % 4 == 2:
            t = t - 3_000_000                                     # events older than the slice start (dvs_flow.h:187-190)
        noise = (rng.random(len(fx)) < 0.07).astype(np.uint8) if k % 3 == 1 else None
        init = None
        if k % 4 == 3:                                            # warm start: the 11 ObjectModel scalars, centre as stored
            init = np.zeros(11)
            init[0], init[1] = rows * 0.5 + rng.uniform(-5, 5), cols * 0.5 + rng.uniform(-5, 5)
            init[7], init[8] = rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05)
            init[9], init[10] = rng.uniform(-1e-4, 1e-4), rng.uniform(-1e-4, 1e-4)
        scale = (1, 3, 5)[k % 3]
        mi = (-1, 10, 1, 25)[k % 4]
        kw = dict(scale=scale, max_iter=mi, init_model=init, noise=noise, rows=rows, cols=cols, want_events=True)
        a = ref.minimize(fx, fy, t, **kw)
        b = oracle_port.minimize(fx, fy, t, **kw)
        tag = "case %d (%dx%d scale %d max_iter %d n %d)" % (k, rows, cols, scale, mi, len(fx))
        assert a["rc"] == b["rc"] and a["iters"] == b["iters"], tag
        assert np.array_equal(a["model"], b["model"]), tag
        assert np.array_equal(a["dividers"], b["dividers"]), tag
        for key in ("pr_x", "pr_y", "nx", "ny"):
            assert np.array_equal(a[key], b[key]), (tag, key)
        seen_rc.add(a["rc"])
    assert seen_rc == {0, 1}                                      # optimised and skipped slices both occurred
