"""Pins the OptimizerLocal oracle: the C restatement in oracle/bf_oracle.c against
(a) fixtures produced by the REAL cv2.GaussianBlur, (b) golden records minted from the reference's own
OptimizerLocal compiled into oracle/_ref (oracle/make_golden_local.py), and (c) that compiled reference
itself when it is present in the snapshot."""
import hashlib
import json
import os

import numpy as np
import pytest

from better_flow_b200 import synth
from helpers import GOLDEN_DIR, golden


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def local_golden():
    return json.load(open(os.path.join(GOLDEN_DIR, "local.json")))


def local_case_events(case):
    """Inputs of a golden case: golden event arrays, or regenerated from (seed, velocity) and checked by hash."""
    src = case["source"]
    _, ev = golden()
    if src[0] == "golden":
        k = src[1]
        fx, fy, t = ev[k + "_fr_x"], ev[k + "_fr_y"], ev[k + "_t_ns"]
        if len(src) > 2:
            fx, fy, t = fx[:src[2]], fy[:src[2]], t[:src[2]]
    elif src[0] == "synth":
        _, cols, rows, rate, dur, seed, vel, slice_s, idx = src
        st = synth.make_stream(cols, rows, rate, dur, seed=seed, vel=tuple(vel))
        sl = synth.cut_slices(st, slice_s)[idx]
        fx, fy, t = sl.fr_x, sl.fr_y, sl.t_ns
    else:
        rng = np.random.Generator(np.random.PCG64(src[1]))
        n = 4000
        t = np.sort(rng.integers(0, 10_000_000, n)).astype(np.int32)[::-1].copy()
        if src[0] == "synth_box":
            b = src[2]
            fx, fy = rng.integers(80, 80 + b + 1, n).astype(np.uint16), rng.integers(100, 100 + b + 1, n).astype(np.uint16)
        else:
            fx, fy = rng.integers(90, 93, n).astype(np.uint16), rng.integers(0, 240, n).astype(np.uint16)
    assert sha(np.concatenate([fx.astype(np.int64), fy.astype(np.int64), t.astype(np.int64)])) == case["input_sha"]
    return fx, fy, t


def test_blur_restatement_equals_real_cv2_fixtures(oracle_port):
    fx = np.load(os.path.join(GOLDEN_DIR, "local_blur.npz"))
    for i in range(5):
        img = fx["in%d" % i]
        for k in (3, 5):
            assert np.array_equal(oracle_port.gaussian_blur_u8(img, k), fx["out%d_k%d" % (i, k)]), (i, k)
        assert np.array_equal(oracle_port.gaussian_blur_u8(img, 1), img)


def test_cv_shim_blur_equals_real_cv2_fixtures():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not present in this snapshot")
    fx = np.load(os.path.join(GOLDEN_DIR, "local_blur.npz"))
    for i in range(5):
        for k in (3, 5):
            assert np.array_equal(ref.blur(fx["in%d" % i], k), fx["out%d_k%d" % (i, k)]), (i, k)


@pytest.mark.parametrize("name", [c["name"] for c in local_golden()["cases"]])
def test_restatement_equals_golden(oracle_port, name):
    case = next(c for c in local_golden()["cases"] if c["name"] == name)
    fx, fy, t = local_case_events(case)
    r = oracle_port.local_minimize(fx, fy, t, case["scale"], rows=case["rows"], cols=case["cols"], want_image=True,
                                   want_events=True)
    assert r["rc"] == case["rc"]
    want = [float.fromhex(h) for h in case["state"]]
    got = [r[k] for k in ("nx", "ny", "score", "dnx", "dny", "dn_th")]
    assert [float(g).hex() for g in got] == [float(w).hex() for w in want]
    assert (r["img_rows"], r["img_cols"]) == (case["img_rows"], case["img_cols"])
    if case["rc"] == 0:
        if case["steps"] >= 0:
            assert r["steps"] == case["steps"]
        assert sha(r["image"]) == case["image_sha"]
        assert sha(np.concatenate([r["pr_x"], r["pr_y"]])) == case["pr_sha"]


def test_restatement_equals_compiled_reference_on_fresh_clouds(oracle_port):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not present in this snapshot")
    for seed, vel, dur, scale in [(71, (30.0, 10.0), 0.02, 3), (72, (-90.0, 50.0), 0.04, 3), (73, (60.0, 60.0), 0.04, 1), (74, (40.0, -20.0), 0.02, 5)]:
        st = synth.make_stream(240, 180, 1.5e6, dur, seed=seed, vel=vel)
        sl = synth.cut_slices(st, dur)[0]
        a = oracle_port.local_minimize(sl.fr_x, sl.fr_y, sl.t_ns, scale, want_image=True)
        b = ref.local_minimize(sl.fr_x, sl.fr_y, sl.t_ns, scale, want_image=True)
        for k in ("rc", "nx", "ny", "score", "dnx", "dny", "dn_th"):
            assert a[k] == b[k], (seed, k)
        if scale > 1:
            assert a["steps"] == b["steps"]
        assert np.array_equal(a["image"], b["image"])
