"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libbf_ref_<rows>x<cols>.so.

That library is the reference's own, unmodified C++ (compiled from /root/reference by
oracle/Makefile); see oracle/ref_driver.cpp for what each entry point wraps.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

MODEL_FIELDS = ("cx", "cy", "dx", "dy", "rot", "div", "cnt",
                "total_dx", "total_dy", "total_rot", "total_div")

_libs: dict[tuple[int, int], C.CDLL] = {}


def lib_path(rows: int, cols: int) -> str:
    return os.path.join(REF_DIR, "libbf_ref_%dx%d.so" % (rows, cols))


def available(rows: int = 180, cols: int = 240) -> bool:
    return os.path.exists(lib_path(rows, cols))


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


def load(rows: int = 180, cols: int = 240) -> C.CDLL:
    key = (rows, cols)
    if key not in _libs:
        lib = C.CDLL(lib_path(rows, cols))
        lib.bf_ref_minimize.restype = C.c_int
        lib.bf_ref_stream.restype = C.c_int
        rx, ry, th = C.c_int(), C.c_int(), C.c_int()
        lib.bf_ref_info(C.byref(rx), C.byref(ry), C.byref(th))
        assert (rx.value, ry.value) == (rows, cols), "library built for another sensor"
        lib.threads = th.value
        _libs[key] = lib
    return _libs[key]


def minimize(fr_x, fr_y, t_ns, scale=3, max_iter=-1, init_model=None, noise=None,
             rows=180, cols=240, want_events=False, slice_start=0):
    """OptimizerRolling on one slice.  ``t_ns`` are local times; because the reference takes
    absolute unsigned timestamps plus a slice start, negative local times are realised by
    shifting both (set_local_time only ever uses the difference, event.h:61-63)."""
    lib = load(rows, cols)
    n = int(len(fr_x))
    fx = np.ascontiguousarray(fr_x, dtype=np.uint32)
    fy = np.ascontiguousarray(fr_y, dtype=np.uint32)
    t = np.asarray(t_ns, dtype=np.int64)
    shift = int(max(0, -int(t.min()))) if n else 0
    ts = np.ascontiguousarray(t + shift + slice_start, dtype=np.uint64)
    start = np.uint64(shift + slice_start)
    nz = np.ascontiguousarray(noise, dtype=np.uint8) if noise is not None else None
    im = np.ascontiguousarray(init_model, dtype=np.float64) if init_model is not None else None
    out_model = np.zeros(11)
    iters = C.c_int(0)
    si = np.zeros(8, dtype=np.int32)
    sd = np.zeros(2)
    div = np.zeros(4, dtype=np.float32)
    pr = np.zeros(4 * n) if want_events else None
    out_noise = np.zeros(n, dtype=np.uint8)
    secs = C.c_double(0)
    rc = lib.bf_ref_minimize(
        C.c_int(n), _p(fx, C.c_uint32), _p(fy, C.c_uint32), _p(ts, C.c_uint64), _p(nz, C.c_uint8),
        C.c_uint64(int(start)), C.c_int(scale), C.c_int(max_iter), _p(im, C.c_double),
        _p(out_model, C.c_double), C.byref(iters), _p(si, C.c_int), _p(sd, C.c_double),
        _p(div, C.c_float), _p(pr, C.c_double), _p(out_noise, C.c_uint8), C.byref(secs))
    res = {
        "rc": rc, "iters": iters.value, "model": out_model, "seconds": secs.value,
        "x_min": int(si[0]), "x_max": int(si[1]), "y_min": int(si[2]), "y_max": int(si[3]),
        "wsize_x": int(si[4]), "wsize_y": int(si[5]), "img_rows": int(si[6]), "img_cols": int(si[7]),
        "x_shift": float(sd[0]), "y_shift": float(sd[1]), "dividers": div, "noise": out_noise,
    }
    if want_events:
        res["pr_x"], res["pr_y"], res["nx"], res["ny"] = pr[:n], pr[n:2 * n], pr[2 * n:3 * n], pr[3 * n:]
    return res


def time_img(pr_x, pr_y, t_local, w, h, scale, x_sh, y_sh, noise=None, rows=180, cols=240):
    lib = load(rows, cols)
    n = int(len(pr_x))
    px = np.ascontiguousarray(pr_x, dtype=np.float64)
    py = np.ascontiguousarray(pr_y, dtype=np.float64)
    t = np.ascontiguousarray(t_local, dtype=np.int64)
    nz = np.ascontiguousarray(noise, dtype=np.uint8) if noise is not None else None
    out = np.zeros((w + scale, h + scale), dtype=np.float32)
    lib.bf_ref_time_img(C.c_int(n), _p(px, C.c_double), _p(py, C.c_double), _p(t, C.c_int64),
                        _p(nz, C.c_uint8), C.c_int(w), C.c_int(h), C.c_int(scale),
                        C.c_int(int(x_sh)), C.c_int(int(y_sh)), _p(out, C.c_float))
    return out


def model(img, want_grad=False, rows=180, cols=240):
    lib = load(rows, cols)
    im = np.ascontiguousarray(img, dtype=np.float32)
    out7 = np.zeros(7)
    gx = np.zeros_like(im) if want_grad else None
    gy = np.zeros_like(im) if want_grad else None
    lib.bf_ref_model(C.c_int(im.shape[0]), C.c_int(im.shape[1]), _p(im, C.c_float),
                     _p(out7, C.c_double), _p(gx, C.c_float), _p(gy, C.c_float))
    return (out7, gx, gy) if want_grad else out7


def project(fr_x, fr_y, t_local, pr_x, pr_y, dnx, dny, cx, cy, div, crl, rows=180, cols=240):
    lib = load(rows, cols)
    n = int(len(fr_x))
    fx = np.ascontiguousarray(fr_x, dtype=np.uint32)
    fy = np.ascontiguousarray(fr_y, dtype=np.uint32)
    t = np.ascontiguousarray(t_local, dtype=np.int64)
    px = np.array(pr_x, dtype=np.float64)
    py = np.array(pr_y, dtype=np.float64)
    nx = np.zeros(n)
    ny = np.zeros(n)
    lib.bf_ref_project(C.c_int(n), _p(fx, C.c_uint32), _p(fy, C.c_uint32), _p(t, C.c_int64),
                       _p(px, C.c_double), _p(py, C.c_double), _p(nx, C.c_double), _p(ny, C.c_double),
                       C.c_double(dnx), C.c_double(dny), C.c_double(cx), C.c_double(cy),
                       C.c_double(div), C.c_double(crl))
    return px, py, nx, ny


def compute_uv(nx, ny, rows=180, cols=240):
    lib = load(rows, cols)
    a = np.ascontiguousarray(nx, dtype=np.float64)
    b = np.ascontiguousarray(ny, dtype=np.float64)
    u = np.zeros_like(a)
    v = np.zeros_like(b)
    lib.bf_ref_compute_uv(C.c_int(len(a)), _p(a, C.c_double), _p(b, C.c_double),
                          _p(u, C.c_double), _p(v, C.c_double))
    return u, v


def stream(fr_x, fr_y, ts_ns, config=0, ev_refresh=20000, time_refresh_ns=33000000, scale=3,
           max_iter=-1, stm_disable=False, flush=True, max_slices=4096, rows=180, cols=240):
    """DVS_flow::add_event over a whole stream; returns (models[k,11], info[k,3])."""
    lib = load(rows, cols)
    n = int(len(fr_x))
    fx = np.ascontiguousarray(fr_x, dtype=np.uint32)
    fy = np.ascontiguousarray(fr_y, dtype=np.uint32)
    ts = np.ascontiguousarray(ts_ns, dtype=np.uint64)
    models = np.zeros((max_slices, 11))
    info = np.zeros((max_slices, 3), dtype=np.int64)
    k = lib.bf_ref_stream(C.c_int(config), C.c_int(n), _p(fx, C.c_uint32), _p(fy, C.c_uint32),
                          _p(ts, C.c_uint64), C.c_ulonglong(ev_refresh), C.c_ulonglong(time_refresh_ns),
                          C.c_int(scale), C.c_int(max_iter), C.c_int(1 if stm_disable else 0),
                          C.c_int(1 if flush else 0), C.c_int(max_slices),
                          _p(models, C.c_double), _p(info, C.c_longlong))
    k = min(k, max_slices)
    return models[:k], info[:k]


def blur(img, ksize, rows=180, cols=240):
    """The cv shim's GaussianBlur stand-in (what the compiled reference's OptimizerLocal calls)."""
    lib = load(rows, cols)
    a = np.array(img, dtype=np.uint8, order="C")
    lib.bf_ref_blur(C.c_int(a.shape[0]), C.c_int(a.shape[1]), C.c_int(ksize), _p(a, C.c_uint8))
    return a


def local_minimize(fr_x, fr_y, t_ns, scale=3, rows=180, cols=240, want_image=False, want_events=False):
    """The reference's own OptimizerLocal(LinearEventCloud*, scale)::run() (optimizer_sampler.cpp:4-38)."""
    lib = load(rows, cols)
    lib.bf_ref_local.restype = C.c_int
    n = int(len(fr_x))
    fx = np.ascontiguousarray(fr_x, dtype=np.uint32)
    fy = np.ascontiguousarray(fr_y, dtype=np.uint32)
    t = np.ascontiguousarray(t_ns, dtype=np.int64)
    out10 = np.zeros(10)
    steps = C.c_int(0)
    secs = C.c_double(0)
    img = None
    if want_image and n:
        r = scale * (int(fx.max()) - int(fx.min())) + scale
        c = scale * (int(fy.max()) - int(fy.min())) + scale
        img = np.zeros((r, c), dtype=np.uint8)
    pr = np.zeros(2 * n) if want_events else None
    rc = lib.bf_ref_local(C.c_int(n), _p(fx, C.c_uint32), _p(fy, C.c_uint32), _p(t, C.c_int64), C.c_int(scale),
                          _p(out10, C.c_double), C.byref(steps), _p(img, C.c_uint8), _p(pr, C.c_double), C.byref(secs))
    res = {"rc": rc, "nx": out10[0], "ny": out10[1], "score": out10[2], "dnx": out10[3], "dny": out10[4],
           "dn_th": out10[5], "wsize_x": int(out10[6]), "wsize_y": int(out10[7]), "img_rows": int(out10[8]),
           "img_cols": int(out10[9]), "steps": steps.value, "seconds": secs.value}
    if want_image:
        res["image"] = img
    if want_events:
        res["pr_x"], res["pr_y"] = pr[:n], pr[n:]
    return res
