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
