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
