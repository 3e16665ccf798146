"""TEST INFRASTRUCTURE ONLY: golden fixtures for the debug images.

EventFile::projection_img (reference: better_flow_core/include/better_flow/event_file.h:460-515) restated in numpy
around the REAL OpenCV calls it makes -- cv2.GaussianBlur(img, (s, s), 0, 0) on CV_8UC1 and cv2.convertScaleAbs -- so
the fixtures carry OpenCV's own fixed-point blur and float rounding (the compiled reference in oracle/_ref links a
header shim whose drawing / conversion calls are no-ops, so it cannot produce these images).
Writes tests/golden/projection_img.npz.  Run here (cv2 is in this image, it is not needed on the GPU box)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def projection_img(pr_x, pr_y, noise, scale, res_x, res_y):
    """event_file.h:460-515 with show_final handled by the caller (pass fr as pr).  Returns (image, nonzero average)."""
    import cv2
    rows, cols = res_x * scale, res_y * scale
    h = scale // 2
    x = np.trunc(np.asarray(pr_x, dtype=np.float64) * scale)            # int x = e.pr_x * scale  (C truncation toward zero)
    y = np.trunc(np.asarray(pr_y, dtype=np.float64) * scale)
    keep = (np.asarray(noise) == 0) & np.isfinite(x) & np.isfinite(y)
    keep &= ~((x >= scale * (res_x - 1)) | (x < 0) | (y >= scale * (res_y - 1)) | (y < 0))      # :486-487
    xi = x[keep].astype(np.int64) + h                                    # :489-490
    yi = y[keep].astype(np.int64) + h
    cnt = np.zeros((rows, cols), dtype=np.int64)
    for dx in range(-h, h + 1):                                          # :494-501 (the clamps never bite: see bf_cuda.h)
        for dy in range(-h, h + 1):
            np.add.at(cnt, (xi + dx, yi + dy), 1)
    img = np.minimum(cnt, 255).astype(np.uint8)                          # "if (< 255) ++" per splat = saturating count
    if scale > 1:
        img = cv2.GaussianBlur(img, (scale, scale), 0, 0)                # :504-506
    nz = img[img != 0]
    avg = float(nz.astype(np.float64).sum() / len(nz)) if len(nz) else 0.0   # EventFile::nonzero_average (event_file.cpp:282-294)
    if avg == 0.0:
        return img, avg
    return cv2.convertScaleAbs(img, alpha=127.0 / avg, beta=0), avg      # :508-509


def color_time_img(pr_x, pr_y, t_local, noise, scale, res_x, res_y):
    """EventFile::color_time_img (event_file.h:649-747) with show_final handled by the caller (pass fr as pr): the
    per-pixel mean direction of the events' local-time "phase" as an HSV image, converted with the real
    cv2.cvtColor(HSV2BGR).  Returns a (scale * res_x + scale, scale * res_y + scale, 3) uint8 BGR image."""
    import cv2
    t = np.asarray(t_local, dtype=np.int64)
    t_min, t_max = int(t.min()), int(t.max())                               # :663-666 (over ALL events, noise included)
    wx, wy = scale * res_x, scale * res_y                                   # :668-680 (the bbox is overridden by the full frame)
    rows, cols = wx + scale, wy + scale
    x_shift = -float((res_x // 2) * scale) + wx / 2.0                       # :690-691 (integer halving of the extent)
    y_shift = -float((res_y // 2) * scale) + wy / 2.0
    x = np.trunc(np.asarray(pr_x, dtype=np.float64) * scale + x_shift)      # :698-699
    y = np.trunc(np.asarray(pr_y, dtype=np.float64) * scale + y_shift)
    keep = (np.asarray(noise) == 0) & ~((x >= wx) | (x < 0) | (y >= wy) | (y < 0))   # :696, :706-709
    ang = (2 * 3.14 * ((t - t_min).astype(np.float64) / float(t_max - t_min))).astype(np.float32)   # :711
    h = scale // 2
    xi, yi, a = x[keep].astype(np.int64) + h, y[keep].astype(np.int64) + h, ang[keep]
    c0 = np.zeros((rows, cols), dtype=np.float32); c1 = np.zeros((rows, cols), dtype=np.float32); cnt = np.zeros((rows, cols), dtype=np.float32)
    co, si = np.cos(a.astype(np.float64)), np.sin(a.astype(np.float64))
    for dx in range(-h, h + 1):                                             # :716-722 (float += double, in event order)
        for dy in range(-h, h + 1):
            np.add.at(c0, (xi + dx, yi + dy), co); np.add.at(c1, (xi + dx, yi + dy), si); np.add.at(cnt, (xi + dx, yi + dy), 1.0)
    hsv = np.zeros((rows, cols, 3), dtype=np.uint8)
    m = cnt >= 1
    vx = np.where(m, c0 / np.maximum(cnt, 1), 0).astype(np.float32); vy = np.where(m, c1 / np.maximum(cnt, 1), 0).astype(np.float32)
    speed = np.hypot(vx.astype(np.float64), vy.astype(np.float64))          # :729-733
    angle = np.where(speed != 0, (np.arctan2(vy.astype(np.float64), vx.astype(np.float64)) + 3.1416) * 180 / 3.1416, 0.0)
    hsv[..., 0] = np.where(m, (angle / 2).astype(np.uint8), 0)              # :735-737 (double -> uchar truncates)
    hsv[..., 1] = np.where(m, np.minimum(speed * 255, 255).astype(np.uint8), 0)
    hsv[..., 2] = np.where(m, 255, 0)
    return cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)                             # :742-746


def cases():
    from better_flow_b200 import synth
    rng = np.random.default_rng(12)
    st = synth.make_stream(240, 180, 2e6, 0.02, seed=21)
    fr_x, fr_y = st.y.astype(np.float64), st.x.astype(np.float64)
    n = len(fr_x)
    warped_x = fr_x + rng.normal(0, 1.2, n) - 0.8          # some events leave the frame (negative / beyond the last row)
    warped_y = fr_y + rng.normal(0, 1.2, n) + 0.6
    noise = (rng.uniform(size=n) < 0.03).astype(np.uint8)
    dense = np.concatenate([np.full(3000, 90.3), warped_x[:5000]]), np.concatenate([np.full(3000, 120.6), warped_y[:5000]])   # a saturating pixel
    return {
        "warped": (warped_x, warped_y, noise), "final": (fr_x, fr_y, noise),
        "dense": (dense[0], dense[1], np.zeros(len(dense[0]), dtype=np.uint8)),
    }


def main():
    out = {}
    for name, (px, py, nz) in cases().items():
        out[name + "_input_checksum"] = np.array([float(np.sum(px * 3 + py)), float(nz.sum())])   # the inputs are regenerated by cases()
        for scale in (1, 3, 5):
            img, avg = projection_img(px, py, nz, scale, 180, 240)
            out["%s_s%d" % (name, scale)] = img
            out["%s_s%d_avg" % (name, scale)] = np.array([avg])
    # colour time images: local times = the events' index order mapped to [0, 20 ms)
    for name, (px, py, nz) in cases().items():
        t = np.linspace(0, 19_999_999, len(px)).astype(np.int64)
        for scale in (1, 3):
            out["%s_color_s%d" % (name, scale)] = color_time_img(px, py, t, nz, scale, 180, 240)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "projection_img.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
