"""The committed golden vectors are what the committed minting scripts produce from the compiled reference today:
re-mint into a scratch directory and compare.  (Needs oracle/_ref, i.e. /root/reference at build time; skipped on
boxes where it did not travel.)"""
import importlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _ref_ready():
    from oracle import ref
    return ref.available(180, 240) and ref.available(260, 346)


def test_golden_json_and_events_are_reproducible(tmp_path, monkeypatch):
    if not _ref_ready():
        pytest.skip("oracle/_ref not built")
    mg = importlib.import_module("oracle.make_golden")
    monkeypatch.setattr(mg, "OUT", str(tmp_path))
    mg.main()
    new = json.load(open(tmp_path / "golden.json"))
    old = json.load(open(os.path.join(GOLD, "golden.json")))
    new.pop("minted_from", None); old.pop("minted_from", None)
    assert new == old
    a, b = np.load(tmp_path / "events.npz"), np.load(os.path.join(GOLD, "events.npz"))
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k


def test_local_goldens_are_reproducible(tmp_path, monkeypatch):
    """OptimizerLocal records (compiled reference class) and the blur fixtures (the real cv2.GaussianBlur)."""
    if not _ref_ready():
        pytest.skip("oracle/_ref not built")
    pytest.importorskip("cv2")
    mg = importlib.import_module("oracle.make_golden_local")
    import shutil
    shutil.copy(os.path.join(GOLD, "events.npz"), tmp_path / "events.npz")   # its input: the committed event fixtures
    monkeypatch.setattr(mg, "GOLD", str(tmp_path))
    mg.main()
    new = json.load(open(tmp_path / "local.json"))
    old = json.load(open(os.path.join(GOLD, "local.json")))
    for d in (new, old):
        for k in [k for k in d if k != "cases"]:
            d.pop(k)                                     # free-text provenance (library versions)
    assert new == old
    a, b = np.load(tmp_path / "local_blur.npz"), np.load(os.path.join(GOLD, "local_blur.npz"))
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k


def test_projection_img_goldens_are_reproducible_and_match_a_plain_restatement():
    """tests/golden/projection_img.npz is what oracle/make_golden_img.py produces with the real cv2 today, and the
    OpenCV calls in it (GaussianBlur on CV_8UC1, convertScaleAbs) equal the plain integer / float32 restatement the CUDA
    path implements: (sum + half) >> shift with binomial weights and reflect-101, rint(float(v) * float(alpha))."""
    pytest.importorskip("cv2")
    mg = importlib.import_module("oracle.make_golden_img")
    from oracle import port
    G = np.load(os.path.join(GOLD, "projection_img.npz"))
    for name, (px, py, nz) in mg.cases().items():
        for scale in (1, 3, 5):
            img, avg = mg.projection_img(px, py, nz, scale, 180, 240)
            assert np.array_equal(img, G["%s_s%d" % (name, scale)]) and avg == G["%s_s%d_avg" % (name, scale)][0]
            # the restatement, without OpenCV
            h = scale // 2
            x, y = np.trunc(px * scale), np.trunc(py * scale)
            keep = (nz == 0) & ~((x >= scale * 179) | (x < 0) | (y >= scale * 239) | (y < 0))
            cnt = np.zeros((180 * scale, 240 * scale), dtype=np.int64)
            for dx in range(-h, h + 1):
                for dy in range(-h, h + 1):
                    np.add.at(cnt, (x[keep].astype(int) + h + dx, y[keep].astype(int) + h + dy), 1)
            a = np.minimum(cnt, 255).astype(np.uint8)
            if scale > 1:
                a = port.gaussian_blur_u8(a, scale)
            nzv = a[a != 0].astype(np.float64)
            alpha = np.float32(127.0 / (nzv.sum() / len(nzv)))
            want = np.minimum(np.rint(np.abs(a.astype(np.float32) * alpha)), 255).astype(np.uint8)
            assert np.array_equal(want, img), (name, scale)
