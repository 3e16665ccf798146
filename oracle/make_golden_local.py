"""TEST INFRASTRUCTURE ONLY.  Mints tests/golden/local.json + local_blur.npz for the OptimizerLocal path
(SURVEY 8a-18 / 8f-3).  Run in the BUILD container (needs /root/reference compiled into oracle/_ref
and the cv2 wheel); the GPU box only reads the committed fixtures.

  * local_blur.npz : random CV_8UC1 images and what the REAL cv2.GaussianBlur(img, (k, k), 0, 0) returns
                     for k = 3, 5 -- pins the blur restatement (oracle/bf_oracle.c, the cv shim);
  * local.json     : OptimizerLocal(LinearEventCloud*, scale)::run() of the reference's own class
                     (oracle/_ref, bf_ref_local) on golden / synthetic clouds: nx, ny, score, dnx, dny as
                     hex floats, step count, SHA-256 of the last image.  Synthetic inputs are regenerated
                     by the tests from (seed, velocity); their SHA-256 guards against generator drift.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from better_flow_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name, source, scale.  source = ("golden", key[, first_n]) or ("synth", cols, rows, rate, dur, seed, vel, slice_s, index)
CASES = [
    ("g240a_s3", ("golden", "davis240_a"), 3),
    ("g240a_s1", ("golden", "davis240_a"), 1),
    ("g240rot_s3", ("golden", "davis240_rot"), 3),
    ("g346_s3", ("golden", "davis346"), 3),
    ("slow_s3", ("synth", 240, 180, 3e6, 0.02, 41, (8.0, -4.0), 0.01, 0), 3),
    ("slow_neg_s3", ("synth", 240, 180, 3e6, 0.02, 42, (-6.0, 7.0), 0.01, 1), 3),
    ("slow_s1", ("synth", 240, 180, 3e6, 0.02, 43, (5.0, 9.0), 0.01, 0), 1),
    ("fast30ms_s3", ("synth", 240, 180, 3e6, 0.03, 47, (80.0, -40.0), 0.03, 0), 3),
    ("fast30ms_s1", ("synth", 240, 180, 2e6, 0.06, 48, (-120.0, 60.0), 0.03, 1), 1),
    ("still_s3", ("synth", 240, 180, 2e6, 0.01, 44, (0.0, 0.0), 0.01, 0), 3),
    ("few_events_s3", ("golden", "davis240_a", 700), 3),       # no 1000-event guard in OptimizerLocal
    ("tiny_window_s3", ("synth_box", 45, 3), 3),                # both image sides below scale*RES/15: run() returns 1
    ("thin_window_s3", ("synth_row", 46), 3),                   # one image side tiny, the other not: runs
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def case_events(src, ev):
    if src[0] == "golden":
        k = src[1]
        fx, fy, t = ev[k + "_fr_x"], ev[k + "_fr_y"], ev[k + "_t_ns"]
        if len(src) > 2:
            fx, fy, t = fx[:src[2]], fy[:src[2]], t[:src[2]]
        return fx, fy, t
    if src[0] == "synth":
        _, cols, rows, rate, dur, seed, vel, slice_s, idx = src
        st = synth.make_stream(cols, rows, rate, dur, seed=seed, vel=vel)
        sl = synth.cut_slices(st, slice_s)[idx]
        return sl.fr_x, sl.fr_y, sl.t_ns
    rng = np.random.Generator(np.random.PCG64(src[1]))
    n = 4000
    t = np.sort(rng.integers(0, 10_000_000, n)).astype(np.int32)[::-1].copy()
    if src[0] == "synth_box":       # events confined to a (box+1)^2 patch
        b = src[2]
        return (rng.integers(80, 80 + b + 1, n).astype(np.uint16), rng.integers(100, 100 + b + 1, n).astype(np.uint16), t)
    # synth_row: rows 90..92, all columns
    return (rng.integers(90, 93, n).astype(np.uint16), rng.integers(0, 240, n).astype(np.uint16), t)


def sensor(src):
    if src[0] == "golden" and src[1] == "davis346":
        return 260, 346
    if src[0] == "synth":
        return src[2], src[1]
    return 180, 240


def main():
    rng = np.random.Generator(np.random.PCG64(2024))
    blur = {}
    for i, (r, c, sparse) in enumerate([(40, 50, False), (37, 64, True), (64, 33, True), (48, 48, False), (90, 120, True)]):
        img = rng.integers(0, 256, (r, c)).astype(np.uint8)
        if sparse:
            img = ((rng.random((r, c)) < 0.15) * rng.integers(1, 256, (r, c))).astype(np.uint8)
        if i == 3:
            img[:] = 255
            img[10:20, 5:9] = 0
        blur["in%d" % i] = img
        for k in (3, 5):
            blur["out%d_k%d" % (i, k)] = cv2.GaussianBlur(img, (k, k), 0, 0)
    np.savez_compressed(os.path.join(GOLD, "local_blur.npz"), **blur)

    ev = np.load(os.path.join(GOLD, "events.npz"))
    out = {"minted_from": "oracle/_ref (reference OptimizerLocal compiled in place, cv::GaussianBlur = oracle/shim restatement "
                          "validated against cv2 %s)" % cv2.__version__, "cases": []}
    for name, src, scale in CASES:
        fx, fy, t = case_events(src, ev)
        rows, cols = sensor(src)
        r = ref.local_minimize(fx, fy, t, scale, rows=rows, cols=cols, want_image=True, want_events=True)
        rec = {"name": name, "source": list(src), "scale": scale, "rows": rows, "cols": cols, "n": int(len(fx)),
               "input_sha": sha(np.concatenate([fx.astype(np.int64), fy.astype(np.int64), t.astype(np.int64)])),
               "rc": r["rc"], "steps": r["steps"],
               "state": [float(r[k]).hex() for k in ("nx", "ny", "score", "dnx", "dny", "dn_th")],
               "img_rows": r["img_rows"], "img_cols": r["img_cols"],
               "image_sha": sha(r["image"]) if r["rc"] == 0 else None,
               "image_nz": int((r["image"] > 0).sum()) if r["rc"] == 0 else 0,
               "pr_sha": sha(np.concatenate([r["pr_x"], r["pr_y"]])) if r["rc"] == 0 else None}
        out["cases"].append(rec)
        print(name, "rc", r["rc"], "steps", r["steps"], "nx %.6g ny %.6g score %.6g" % (r["nx"], r["ny"], r["score"]), "nz", rec["image_nz"])
    json.dump(out, open(os.path.join(GOLD, "local.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
