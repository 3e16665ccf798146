"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/libbf_oracle.so (the C restatement,
oracle/bf_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libbf_oracle.so")

_lib = None


class Setup(C.Structure):
    _fields_ = [("x_min", C.c_int), ("x_max", C.c_int), ("y_min", C.c_int), ("y_max", C.c_int),
                ("wsize_x", C.c_int), ("wsize_y", C.c_int), ("img_rows", C.c_int), ("img_cols", C.c_int),
                ("x_shift", C.c_double), ("y_shift", C.c_double)]


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "bf_oracle.c")):
            build()
        _lib = C.CDLL(LIB)
        _lib.bfo_minimize.restype = C.c_int
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


def setup_slice(fr_x, fr_y, rows, cols, scale):
    lib = load()
    fx = np.ascontiguousarray(fr_x, dtype=np.uint16)
    fy = np.ascontiguousarray(fr_y, dtype=np.uint16)
    s = Setup()
    lib.bfo_setup_slice(C.c_int(len(fx)), _p(fx, C.c_uint16), _p(fy, C.c_uint16), C.c_int(rows),
                        C.c_int(cols), C.c_int(scale), C.byref(s))
    return s


def time_img(pr_x, pr_y, t_local, w, h, scale, x_sh, y_sh, noise=None, accum_mode=0):
    lib = load()
    px = np.ascontiguousarray(pr_x, dtype=np.float64)
    py = np.ascontiguousarray(pr_y, dtype=np.float64)
    t = np.ascontiguousarray(t_local, dtype=np.int64)
    nz = np.ascontiguousarray(noise, dtype=np.uint8) if noise is not None else None
    out = np.zeros((w + scale, h + scale), dtype=np.float32)
    lib.bfo_time_img(C.c_int(len(px)), _p(px, C.c_double), _p(py, C.c_double), _p(t, C.c_int64),
                     _p(nz, C.c_uint8), C.c_int(w), C.c_int(h), C.c_int(scale), C.c_int(int(x_sh)),
                     C.c_int(int(y_sh)), C.c_int(accum_mode), _p(out, C.c_float))
    return out


def model(img, want_grad=False):
    lib = load()
    im = np.ascontiguousarray(img, dtype=np.float32)
    out7 = np.zeros(7)
    gx = np.zeros_like(im)
    gy = np.zeros_like(im)
    lib.bfo_model_update(C.c_int(im.shape[0]), C.c_int(im.shape[1]), _p(im, C.c_float),
                         _p(out7, C.c_double), _p(gx, C.c_float), _p(gy, C.c_float))
    return (out7, gx, gy) if want_grad else out7


def project(fr_x, fr_y, t_local, pr_x, pr_y, dnx, dny, cx, cy, div, crl):
    lib = load()
    n = len(fr_x)
    fx = np.ascontiguousarray(fr_x, dtype=np.uint16)
    fy = np.ascontiguousarray(fr_y, dtype=np.uint16)
    t = np.ascontiguousarray(t_local, dtype=np.int64)
    px = np.array(pr_x, dtype=np.float64)
    py = np.array(pr_y, dtype=np.float64)
    nx = np.zeros(n)
    ny = np.zeros(n)
    lib.bfo_project(C.c_int(n), _p(fx, C.c_uint16), _p(fy, C.c_uint16), _p(t, C.c_int64),
                    _p(px, C.c_double), _p(py, C.c_double), _p(nx, C.c_double), _p(ny, C.c_double),
                    C.c_double(dnx), C.c_double(dny), C.c_double(cx), C.c_double(cy),
                    C.c_double(div), C.c_double(crl))
    return px, py, nx, ny


def compute_uv(nx, ny):
    lib = load()
    a = np.ascontiguousarray(nx, dtype=np.float64)
    b = np.ascontiguousarray(ny, dtype=np.float64)
    u = np.zeros_like(a)
    v = np.zeros_like(b)
    lib.bfo_compute_uv(C.c_int(len(a)), _p(a, C.c_double), _p(b, C.c_double), _p(u, C.c_double), _p(v, C.c_double))
    return u, v


def minimize(fr_x, fr_y, t_ns, scale=3, max_iter=-1, init_model=None, noise=None,
             rows=180, cols=240, accum_mode=0, want_events=False):
    lib = load()
    n = int(len(fr_x))
    fx = np.ascontiguousarray(fr_x, dtype=np.uint16)
    fy = np.ascontiguousarray(fr_y, dtype=np.uint16)
    t = np.ascontiguousarray(t_ns, dtype=np.int64)
    nz = np.array(noise, dtype=np.uint8) if noise is not None else np.zeros(n, dtype=np.uint8)
    im = np.ascontiguousarray(init_model, dtype=np.float64) if init_model is not None else None
    out_model = np.zeros(11)
    iters = C.c_int(0)
    su = Setup()
    div = np.zeros(4, dtype=np.float32)
    pr = np.zeros(4 * n) if want_events else None
    rc = lib.bfo_minimize(C.c_int(n), _p(fx, C.c_uint16), _p(fy, C.c_uint16), _p(t, C.c_int64),
                          _p(nz, C.c_uint8), C.c_int(rows), C.c_int(cols), C.c_int(scale),
                          C.c_int(max_iter), _p(im, C.c_double), C.c_int(accum_mode),
                          _p(out_model, C.c_double), C.byref(iters), C.byref(su), _p(div, C.c_float),
                          _p(pr, C.c_double))
    res = {
        "rc": rc, "iters": iters.value, "model": out_model,
        "x_min": su.x_min, "x_max": su.x_max, "y_min": su.y_min, "y_max": su.y_max,
        "wsize_x": su.wsize_x, "wsize_y": su.wsize_y, "img_rows": su.img_rows, "img_cols": su.img_cols,
        "x_shift": su.x_shift, "y_shift": su.y_shift, "dividers": div, "noise": nz,
    }
    if want_events:
        res["pr_x"], res["pr_y"], res["nx"], res["ny"] = pr[:n], pr[n:2 * n], pr[2 * n:3 * n], pr[3 * n:]
    return res


def gaussian_blur_u8(img, ksize):
    """bfo_gaussian_blur_u8: the restated cv::GaussianBlur(img, img, Size(k, k), 0, 0) for CV_8UC1."""
    lib = load()
    a = np.array(img, dtype=np.uint8, order="C")
    lib.bfo_gaussian_blur_u8(C.c_int(a.shape[0]), C.c_int(a.shape[1]), C.c_int(ksize), _p(a, C.c_uint8))
    return a


def local_minimize(fr_x, fr_y, t_ns, scale=3, rows=180, cols=240, want_image=False, want_events=False):
    """OptimizerLocal(cloud, scale).run() restated (optimizer_sampler.cpp)."""
    lib = load()
    lib.bfo_local_minimize.restype = C.c_int
    n = int(len(fr_x))
    fx = np.ascontiguousarray(fr_x, dtype=np.uint16)
    fy = np.ascontiguousarray(fr_y, dtype=np.uint16)
    t = np.ascontiguousarray(t_ns, dtype=np.int64)
    out10 = np.zeros(10)
    steps = C.c_int(0)
    su = setup_slice(fx, fy, rows, cols, scale) if n else None
    img = np.zeros((su.img_rows, su.img_cols), dtype=np.uint8) if (want_image and n) else None
    pr = np.zeros(2 * n) if want_events else None
    rc = lib.bfo_local_minimize(C.c_int(n), _p(fx, C.c_uint16), _p(fy, C.c_uint16), _p(t, C.c_int64), C.c_int(rows),
                                C.c_int(cols), C.c_int(scale), _p(out10, C.c_double), C.byref(steps),
                                _p(img, C.c_uint8), _p(pr, C.c_double))
    res = {"rc": rc, "nx": out10[0], "ny": out10[1], "score": out10[2], "dnx": out10[3], "dny": out10[4],
           "dn_th": out10[5], "wsize_x": int(out10[6]), "wsize_y": int(out10[7]), "img_rows": int(out10[8]),
           "img_cols": int(out10[9]), "steps": steps.value}
    if want_image:
        res["image"] = img
    if want_events:
        res["pr_x"], res["pr_y"] = pr[:n], pr[n:]
    return res
