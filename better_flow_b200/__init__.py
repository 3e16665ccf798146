"""better_flow_b200 -- B200-native motion compensation for better-flow.

The product is the C-ABI shared library ``libbf_cuda.so`` (include/bf_cuda.h) plus the C++ host
mirror of the reference classes under ``include/better_flow``.  This Python module is only a thin
ctypes binding of that C ABI, used by the tests and by bench.py; it contains no compute and no
fallback: if the library or a CUDA device is missing, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# BF_LIB_PATH: development aid for same-box A/B runs of two builds of the library (tools/ab_libs.py)
LIB_PATH = os.environ.get("BF_LIB_PATH") or os.path.join(HERE, "libbf_cuda.so")

RESULT_BYTES = 160  # sizeof(bf_slice_result), checked in load()
RC_OK, RC_SKIPPED, RC_ITER_CAP, RC_DEGENERATE = 0, 1, 2, 3
FLAG_ALL_NOISE, FLAG_T_QUANTISED = 1, 2
EVENT_NOISE = 0x8000

EVENT_DTYPE = np.dtype([("fr_x", "<u2"), ("fr_y", "<u2"), ("t_ns", "<i4")])
RING_EVENT_DTYPE = np.dtype([("fr_x", "<u2"), ("fr_y", "<u2"), ("reserved", "<u4"), ("timestamp", "<u8")])   # bf_ring_event


class Model(C.Structure):
    _fields_ = [("cx", C.c_double), ("cy", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
                ("rot", C.c_double), ("div", C.c_double), ("cnt", C.c_uint32), ("pad_", C.c_uint32),
                ("total_dx", C.c_double), ("total_dy", C.c_double), ("total_rot", C.c_double),
                ("total_div", C.c_double)]

    def as_array(self) -> np.ndarray:
        return np.array([self.cx, self.cy, self.dx, self.dy, self.rot, self.div, float(self.cnt),
                         self.total_dx, self.total_dy, self.total_rot, self.total_div])

    @classmethod
    def from_array(cls, a) -> "Model":
        m = cls()
        (m.cx, m.cy, m.dx, m.dy, m.rot, m.div) = [float(v) for v in a[:6]]
        m.cnt = int(a[6])
        (m.total_dx, m.total_dy, m.total_rot, m.total_div) = [float(v) for v in a[7:11]]
        return m


class SliceResult(C.Structure):
    _fields_ = [("model", Model), ("rc", C.c_int32), ("iters", C.c_int32), ("dividers", C.c_float * 4),
                ("x_min", C.c_int32), ("x_max", C.c_int32), ("y_min", C.c_int32), ("y_max", C.c_int32),
                ("img_rows", C.c_int32), ("img_cols", C.c_int32), ("x_shift", C.c_double),
                ("y_shift", C.c_double), ("n_events", C.c_int32), ("flags", C.c_uint32)]


# every symbol include/bf_cuda.h declares
ABI_SYMBOLS = (
    "bf_cuda_init", "bf_device_count", "bf_last_error", "bf_version", "bf_ctx_create", "bf_ctx_destroy",
    "bf_ctx_set_option", "bf_ctx_get_option", "bf_batch_reset", "bf_batch_add", "bf_batch_add_packed",
    "bf_batch_staging", "bf_batch_add_staged", "bf_batch_upload", "bf_batch_launch", "bf_batch_download",
    "bf_batch_sync", "bf_batch_run", "bf_batch_time_launches", "bf_ctx_launch_count", "bf_batch_size",
    "bf_batch_result", "bf_batch_events", "bf_minimize", "bf_time_img", "bf_fast_model", "bf_project",
    "bf_ctx_set_stream", "bf_batch_results_device", "bf_debug_profile", "bf_model_from_image",
    "bf_batch_run_streamed",
    "bf_batch_add_local", "bf_batch_slot_mode", "bf_local_minimize",
    "bf_multi_create", "bf_multi_destroy", "bf_multi_device_count", "bf_multi_set_option", "bf_multi_owner",
    "bf_multi_reset", "bf_multi_add_packed", "bf_multi_run", "bf_multi_sync", "bf_multi_size", "bf_multi_result",
    "bf_multi_locate", "bf_multi_launch_count",
    "bf_projection_img", "bf_color_time_img", "bf_batch_add_delta", "bf_batch_upload_bytes",
    "bf_ring_create", "bf_ring_destroy", "bf_ring_push", "bf_ring_slice", "bf_ring_result", "bf_ring_sync", "bf_ring_pushed", "bf_ring_seed", "bf_ring_reserve", "bf_ring_commit",
)

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libbf_cuda.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", HERE, "lib"] + ([] if verbose else ["-s"]))
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libbf_cuda.so is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no fallback path)")
        lib = C.CDLL(LIB_PATH)
        lib.bf_last_error.restype = C.c_char_p
        lib.bf_version.restype = C.c_char_p
        lib.bf_ctx_create.restype = C.c_void_p
        lib.bf_ctx_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int]
        lib.bf_ctx_destroy.argtypes = [C.c_void_p]
        lib.bf_ctx_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
        lib.bf_ctx_get_option.argtypes = [C.c_void_p, C.c_char_p]
        lib.bf_ctx_get_option.restype = C.c_longlong
        lib.bf_ctx_launch_count.argtypes = [C.c_void_p]
        lib.bf_ctx_launch_count.restype = C.c_longlong
        lib.bf_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        lib.bf_batch_results_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]
        lib.bf_batch_staging.restype = C.c_void_p
        lib.bf_batch_staging.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        lib.bf_batch_add_staged.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p]
        for name in ("bf_batch_reset", "bf_batch_upload", "bf_batch_download", "bf_batch_sync", "bf_batch_size"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.bf_batch_launch.argtypes = [C.c_void_p, C.c_int]
        lib.bf_batch_run.argtypes = [C.c_void_p, C.c_int]
        lib.bf_batch_run_streamed.argtypes = [C.c_void_p, C.c_int]
        lib.bf_batch_time_launches.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
        lib.bf_batch_result.argtypes = [C.c_void_p, C.c_int, C.POINTER(SliceResult)]
        lib.bf_batch_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p]
        lib.bf_batch_add_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.bf_batch_events.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.bf_minimize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.POINTER(SliceResult), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
        lib.bf_time_img.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.bf_fast_model.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.bf_project.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.c_double, C.c_double]
        lib.bf_batch_add_local.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.bf_batch_slot_mode.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.bf_local_minimize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(SliceResult)]
        lib.bf_multi_create.restype = C.c_void_p
        lib.bf_multi_create.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int]
        for name in ("bf_multi_destroy", "bf_multi_device_count", "bf_multi_reset", "bf_multi_sync", "bf_multi_size"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.bf_multi_destroy.restype = None
        lib.bf_multi_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
        lib.bf_multi_owner.argtypes = [C.c_int, C.c_int, C.c_int]
        lib.bf_multi_add_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.bf_multi_run.argtypes = [C.c_void_p, C.c_int]
        lib.bf_multi_result.argtypes = [C.c_void_p, C.c_int, C.POINTER(SliceResult)]
        lib.bf_multi_locate.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.bf_multi_launch_count.argtypes = [C.c_void_p]
        lib.bf_multi_launch_count.restype = C.c_longlong
        lib.bf_batch_add_delta.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.bf_batch_upload_bytes.argtypes = [C.c_void_p]
        lib.bf_batch_upload_bytes.restype = C.c_longlong
        lib.bf_color_time_img.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.bf_projection_img.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.bf_ring_create.restype = C.c_void_p
        lib.bf_ring_create.argtypes = [C.c_void_p, C.c_longlong, C.c_int]
        lib.bf_ring_destroy.argtypes = [C.c_void_p]
        lib.bf_ring_destroy.restype = None
        lib.bf_ring_push.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.bf_ring_slice.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int]
        lib.bf_ring_result.argtypes = [C.c_void_p, C.c_int, C.POINTER(SliceResult)]
        lib.bf_ring_sync.argtypes = [C.c_void_p]
        lib.bf_ring_reserve.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        lib.bf_ring_commit.argtypes = [C.c_void_p, C.c_int]
        lib.bf_ring_seed.argtypes = [C.c_void_p, C.POINTER(Model)]
        lib.bf_ring_pushed.argtypes = [C.c_void_p]
        lib.bf_ring_pushed.restype = C.c_longlong
        assert C.sizeof(SliceResult) == RESULT_BYTES and C.sizeof(Model) == 88
        _lib = lib
    return _lib


class BfError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def pack_events(fr_x, fr_y, t_ns, noise=None) -> np.ndarray:
    """SoA -> the 8-byte bf_event records (numpy structured array)."""
    n = len(fr_x)
    ev = np.empty(n, dtype=EVENT_DTYPE)
    ev["fr_x"] = fr_x
    fy = np.asarray(fr_y, dtype=np.uint16)
    if noise is not None:
        fy = fy | (np.asarray(noise, dtype=np.uint16) != 0).astype(np.uint16) * np.uint16(EVENT_NOISE)
    ev["fr_y"] = fy
    ev["t_ns"] = t_ns
    return ev


class Context:
    """Owns one bf_ctx (device buffers, stream)."""

    def __init__(self, sensor_rows=180, sensor_cols=240, max_scale=3, max_events=1 << 20, max_slices=64,
                 device=None):
        self.lib = load()
        if device is not None and self.lib.bf_cuda_init(int(device)) != 0:
            raise BfError(self.lib.bf_last_error().decode())
        self.h = self.lib.bf_ctx_create(sensor_rows, sensor_cols, max_scale, int(max_events), int(max_slices))
        if not self.h:
            raise BfError(self.lib.bf_last_error().decode())
        self.rows, self.cols = sensor_rows, sensor_cols

    def close(self):
        if getattr(self, "h", None):
            self.lib.bf_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise BfError(self.lib.bf_last_error().decode())
        return rc

    def set_option(self, key, value):
        self._chk(self.lib.bf_ctx_set_option(self.h, key.encode(), int(value)))

    def get_option(self, key):
        return int(self.lib.bf_ctx_get_option(self.h, key.encode()))

    def set_stream(self, cuda_stream_handle):
        """Run on a caller-owned CUDA stream (integer handle, e.g. torch.cuda.Stream().cuda_stream)."""
        self._chk(self.lib.bf_ctx_set_stream(self.h, C.c_void_p(cuda_stream_handle) if cuda_stream_handle else None))

    def results_device(self):
        """(device pointer, byte size) of the bf_slice_result records of the current batch."""
        p = C.c_void_p(0)
        n = C.c_longlong(0)
        self._chk(self.lib.bf_batch_results_device(self.h, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def debug_profile(self):
        out = np.zeros((1024, 16), dtype=np.int64)
        self.lib.bf_debug_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        n = self._chk(self.lib.bf_debug_profile(self.h, _ptr(out), 1024))
        return out[:n]

    @property
    def launches(self):
        return int(self.lib.bf_ctx_launch_count(self.h))

    # ---- batch ----
    def reset(self):
        self._chk(self.lib.bf_batch_reset(self.h))

    def add(self, fr_x, fr_y, t_ns, scale=3, max_iter=-1, init=None, noise=None):
        fx = np.ascontiguousarray(fr_x, dtype=np.uint16)
        fy = np.ascontiguousarray(fr_y, dtype=np.uint16)
        t = np.ascontiguousarray(t_ns, dtype=np.int32)
        nz = np.ascontiguousarray(noise, dtype=np.uint8) if noise is not None else None
        m = Model.from_array(init) if init is not None else None
        return self._chk(self.lib.bf_batch_add(self.h, _ptr(fx), _ptr(fy), _ptr(t), _ptr(nz), len(fx), scale,
                                               max_iter, C.byref(m) if m is not None else None))

    def add_packed(self, events, scale=3, max_iter=-1, init=None):
        ev = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        m = Model.from_array(init) if init is not None else None
        return self._chk(self.lib.bf_batch_add_packed(self.h, _ptr(ev), len(ev), scale, max_iter,
                                                      C.byref(m) if m is not None else None))

    def add_delta(self, events, scale=3, max_iter=-1, init=None):
        """Queue a slice in the compact 6-byte upload format (raises BfError when it cannot be represented)."""
        ev = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        m = Model.from_array(init) if init is not None else None
        return self._chk(self.lib.bf_batch_add_delta(self.h, _ptr(ev), len(ev), scale, max_iter,
                                                     C.byref(m) if m is not None else None))

    @property
    def upload_bytes(self):
        return int(self.lib.bf_batch_upload_bytes(self.h))

    def staging(self) -> np.ndarray:
        cap = C.c_longlong(0)
        p = self.lib.bf_batch_staging(self.h, C.byref(cap))
        buf = (C.c_char * (cap.value * 8)).from_address(p)
        return np.frombuffer(buf, dtype=EVENT_DTYPE)

    def add_staged(self, offset, n, scale=3, max_iter=-1, init=None):
        m = Model.from_array(init) if init is not None else None
        return self._chk(self.lib.bf_batch_add_staged(self.h, int(offset), int(n), scale, max_iter,
                                                      C.byref(m) if m is not None else None))

    def add_local(self, fr_x, fr_y, t_ns, scale=3):
        """Queue a cloud for OptimizerLocal::run (contrast-driven nx, ny descent)."""
        fx = np.ascontiguousarray(fr_x, dtype=np.uint16)
        fy = np.ascontiguousarray(fr_y, dtype=np.uint16)
        t = np.ascontiguousarray(t_ns, dtype=np.int32)
        return self._chk(self.lib.bf_batch_add_local(self.h, _ptr(fx), _ptr(fy), _ptr(t), len(fx), scale))

    def slot_mode(self, slot, mode):
        self._chk(self.lib.bf_batch_slot_mode(self.h, slot, mode))

    @staticmethod
    def local_view(res: dict) -> dict:
        """Names for the fields of an OptimizerLocal result record (see include/bf_cuda.h)."""
        m = res["model"]
        return {"rc": res["rc"], "steps": res["iters"], "nx": m[7], "ny": m[8], "score": m[2], "dnx": m[3], "dny": m[4],
                "dn_th": m[5], "nz_cnt": int(m[6]), "img_rows": res["img_rows"], "img_cols": res["img_cols"],
                "n_events": res["n_events"]}

    def local_minimize(self, fr_x, fr_y, t_ns, scale=3):
        self.reset()
        self.add_local(fr_x, fr_y, t_ns, scale)
        self.run()
        return self.local_view(self.result(0))

    def upload(self):
        self._chk(self.lib.bf_batch_upload(self.h))

    def launch(self, want_events=False):
        self._chk(self.lib.bf_batch_launch(self.h, 1 if want_events else 0))

    def download(self):
        self._chk(self.lib.bf_batch_download(self.h))

    def sync(self):
        self._chk(self.lib.bf_batch_sync(self.h))

    def run(self, want_events=False):
        self._chk(self.lib.bf_batch_run(self.h, 1 if want_events else 0))

    def run_streamed(self, want_events=False):
        """Asynchronous upload (streamed, overlapped) + launch + download; call sync() afterwards."""
        self._chk(self.lib.bf_batch_run_streamed(self.h, 1 if want_events else 0))

    def time_launches(self, reps, want_events=False) -> float:
        ms = C.c_float(0)
        self._chk(self.lib.bf_batch_time_launches(self.h, reps, 1 if want_events else 0, C.byref(ms)))
        return float(ms.value)

    def size(self):
        return int(self.lib.bf_batch_size(self.h))

    def result(self, slot) -> dict:
        r = SliceResult()
        self._chk(self.lib.bf_batch_result(self.h, slot, C.byref(r)))
        return {
            "rc": r.rc, "iters": r.iters, "model": r.model.as_array(),
            "dividers": np.array(list(r.dividers), dtype=np.float32),
            "x_min": r.x_min, "x_max": r.x_max, "y_min": r.y_min, "y_max": r.y_max,
            "img_rows": r.img_rows, "img_cols": r.img_cols, "x_shift": r.x_shift, "y_shift": r.y_shift,
            "n_events": r.n_events, "flags": r.flags,
        }

    def results(self):
        return [self.result(i) for i in range(self.size())]

    def events(self, slot, n):
        out = [np.zeros(n) for _ in range(4)]
        self._chk(self.lib.bf_batch_events(self.h, slot, *[_ptr(a) for a in out]))
        return dict(zip(("pr_x", "pr_y", "nx", "ny"), out))

    def minimize(self, fr_x, fr_y, t_ns, scale=3, max_iter=-1, init=None, noise=None, want_events=False):
        self.reset()
        self.add(fr_x, fr_y, t_ns, scale, max_iter, init, noise)
        self.run(want_events)
        res = self.result(0)
        if want_events:
            res.update(self.events(0, len(fr_x)))
        return res

    # ---- stage level ----
    def time_img(self, pr_x, pr_y, t_ns, w, h, scale, x_sh, y_sh, noise=None):
        px = np.ascontiguousarray(pr_x, dtype=np.float64)
        py = np.ascontiguousarray(pr_y, dtype=np.float64)
        t = np.ascontiguousarray(t_ns, dtype=np.int32)
        nz = np.ascontiguousarray(noise, dtype=np.uint8) if noise is not None else None
        out = np.zeros((w + scale, h + scale), dtype=np.float32)
        self._chk(self.lib.bf_time_img(self.h, len(px), _ptr(px), _ptr(py), _ptr(t), _ptr(nz), w, h, scale,
                                       int(x_sh), int(y_sh), _ptr(out)))
        return out

    def fast_model(self, pr_x, pr_y, t_ns, w, h, scale, x_sh, y_sh, noise=None, want_grad=False):
        px = np.ascontiguousarray(pr_x, dtype=np.float64)
        py = np.ascontiguousarray(pr_y, dtype=np.float64)
        t = np.ascontiguousarray(t_ns, dtype=np.int32)
        nz = np.ascontiguousarray(noise, dtype=np.uint8) if noise is not None else None
        out7 = np.zeros(7)
        gx = np.zeros((w + scale, h + scale), dtype=np.float32) if want_grad else None
        gy = np.zeros((w + scale, h + scale), dtype=np.float32) if want_grad else None
        self._chk(self.lib.bf_fast_model(self.h, len(px), _ptr(px), _ptr(py), _ptr(t), _ptr(nz), w, h, scale,
                                         int(x_sh), int(y_sh), _ptr(out7), _ptr(gx), _ptr(gy)))
        return (out7, gx, gy) if want_grad else out7

    def model_from_image(self, img, want_grad=False):
        im = np.ascontiguousarray(img, dtype=np.float32)
        out7 = np.zeros(7)
        gx = np.zeros_like(im) if want_grad else None
        gy = np.zeros_like(im) if want_grad else None
        self.lib.bf_model_from_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self._chk(self.lib.bf_model_from_image(self.h, im.shape[0], im.shape[1], _ptr(im), _ptr(out7), _ptr(gx), _ptr(gy)))
        return (out7, gx, gy) if want_grad else out7

    def projection_img(self, pr_x, pr_y, scale=3, noise=None):
        """EventFile::projection_img: (uint8 image [rows * scale, cols * scale], nonzero average before scaling)."""
        px = np.ascontiguousarray(pr_x, dtype=np.float64)
        py = np.ascontiguousarray(pr_y, dtype=np.float64)
        nz = np.ascontiguousarray(noise, dtype=np.uint8) if noise is not None else None
        out = np.zeros((self.rows * scale, self.cols * scale), dtype=np.uint8)
        avg = np.zeros(1)
        self._chk(self.lib.bf_projection_img(self.h, len(px), _ptr(px), _ptr(py), _ptr(nz), scale, _ptr(out), _ptr(avg)))
        return out, float(avg[0])

    def color_time_img(self, pr_x, pr_y, t_ns, scale=3, noise=None):
        """EventFile::color_time_img: uint8 BGR image [rows * scale + scale, cols * scale + scale, 3]."""
        px = np.ascontiguousarray(pr_x, dtype=np.float64)
        py = np.ascontiguousarray(pr_y, dtype=np.float64)
        t = np.ascontiguousarray(t_ns, dtype=np.int32)
        nz = np.ascontiguousarray(noise, dtype=np.uint8) if noise is not None else None
        out = np.zeros((self.rows * scale + scale, self.cols * scale + scale, 3), dtype=np.uint8)
        self._chk(self.lib.bf_color_time_img(self.h, len(px), _ptr(px), _ptr(py), _ptr(t), _ptr(nz), scale, _ptr(out)))
        return out

    def project(self, fr_x, fr_y, t_ns, pr_x, pr_y, dnx, dny, cx, cy, div, crl):
        fx = np.ascontiguousarray(fr_x, dtype=np.uint16)
        fy = np.ascontiguousarray(fr_y, dtype=np.uint16)
        t = np.ascontiguousarray(t_ns, dtype=np.int32)
        px = np.array(pr_x, dtype=np.float64)
        py = np.array(pr_y, dtype=np.float64)
        nx = np.zeros(len(fx))
        ny = np.zeros(len(fx))
        self._chk(self.lib.bf_project(self.h, len(fx), _ptr(fx), _ptr(fy), _ptr(t), _ptr(px), _ptr(py), _ptr(nx),
                                      _ptr(ny), dnx, dny, cx, cy, div, crl))
        return px, py, nx, ny


class Ring:
    """bf_ring: the device-resident slice ring of DVS_flow's default mode (include/bf_cuda.h)."""

    def __init__(self, ctx: "Context", capacity=50000, max_pending=64):
        self.ctx = ctx
        self.lib = ctx.lib
        self.h = self.lib.bf_ring_create(ctx.h, int(capacity), int(max_pending))
        if not self.h:
            raise BfError(self.lib.bf_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.bf_ring_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise BfError(self.lib.bf_last_error().decode())
        return rc

    def push(self, fr_x, fr_y, timestamp_ns):
        ev = np.zeros(len(fr_x), dtype=RING_EVENT_DTYPE)
        ev["fr_x"], ev["fr_y"], ev["timestamp"] = fr_x, fr_y, timestamp_ns
        self._chk(self.lib.bf_ring_push(self.h, _ptr(ev), len(ev)))

    def push_in_place(self, fr_x, fr_y, timestamp_ns, reserve=None):
        """bf_ring_reserve / bf_ring_commit: the events are written straight into the ring's pinned staging buffer."""
        n = len(fr_x)
        where = C.c_void_p()
        self._chk(self.lib.bf_ring_reserve(self.h, int(reserve or max(n, 1)), C.byref(where)))
        ev = np.ctypeslib.as_array(C.cast(where, C.POINTER(C.c_uint8)), shape=(n * RING_EVENT_DTYPE.itemsize,)).view(RING_EVENT_DTYPE) if n else None
        if n:
            ev["fr_x"], ev["fr_y"], ev["reserved"], ev["timestamp"] = fr_x, fr_y, 0, timestamp_ns
        self._chk(self.lib.bf_ring_commit(self.h, n))

    def slice(self, n, slice_start, scale=3, max_iter=-1, chain=True) -> int:
        return self._chk(self.lib.bf_ring_slice(self.h, int(n), int(slice_start), scale, max_iter, 1 if chain else 0))

    def result(self, ticket) -> dict:
        r = SliceResult()
        self._chk(self.lib.bf_ring_result(self.h, int(ticket), C.byref(r)))
        return _result_dict(r)

    def sync(self):
        self._chk(self.lib.bf_ring_sync(self.h))

    def seed(self, model):
        """The model the next chained slice starts from (11 values as Model.as_array returns them)."""
        self._chk(self.lib.bf_ring_seed(self.h, C.byref(Model.from_array(model))))

    @property
    def pushed(self):
        return int(self.lib.bf_ring_pushed(self.h))


def _result_dict(r: SliceResult) -> dict:
    return {
        "rc": r.rc, "iters": r.iters, "model": r.model.as_array(),
        "dividers": np.array(list(r.dividers), dtype=np.float32),
        "x_min": r.x_min, "x_max": r.x_max, "y_min": r.y_min, "y_max": r.y_max,
        "img_rows": r.img_rows, "img_cols": r.img_cols, "x_shift": r.x_shift, "y_shift": r.y_shift,
        "n_events": r.n_events, "flags": r.flags,
    }


class MultiContext:
    """bf_multi: one host process, N devices, slices dealt block-cyclically, one NCCL all-gather of the
    per-slice result records per batch (include/bf_cuda.h)."""

    def __init__(self, n_devices, sensor_rows=180, sensor_cols=240, max_scale=3, max_events_per_device=1 << 20,
                 max_slices_per_device=64, devices=None):
        self.lib = load()
        dev = (C.c_int * n_devices)(*devices) if devices is not None else None
        self.h = self.lib.bf_multi_create(n_devices, dev, sensor_rows, sensor_cols, max_scale,
                                          int(max_events_per_device), int(max_slices_per_device))
        if not self.h:
            raise BfError(self.lib.bf_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.bf_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise BfError(self.lib.bf_last_error().decode())
        return rc

    @property
    def n_devices(self):
        return int(self.lib.bf_multi_device_count(self.h))

    @property
    def launches(self):
        return int(self.lib.bf_multi_launch_count(self.h))

    def set_option(self, key, value):
        self._chk(self.lib.bf_multi_set_option(self.h, key.encode(), int(value)))

    def reset(self):
        self._chk(self.lib.bf_multi_reset(self.h))

    def add_packed(self, events, scale=3, max_iter=-1):
        ev = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        return self._chk(self.lib.bf_multi_add_packed(self.h, _ptr(ev), len(ev), scale, max_iter))

    def run(self, want_events=False):
        self._chk(self.lib.bf_multi_run(self.h, 1 if want_events else 0))

    def sync(self):
        self._chk(self.lib.bf_multi_sync(self.h))

    def size(self):
        return int(self.lib.bf_multi_size(self.h))

    def result(self, k) -> dict:
        r = SliceResult()
        self._chk(self.lib.bf_multi_result(self.h, k, C.byref(r)))
        return _result_dict(r)

    def results(self):
        return [self.result(k) for k in range(self.size())]

    def locate(self, k):
        ctx, slot, dev = C.c_void_p(0), C.c_int(0), C.c_int(0)
        self._chk(self.lib.bf_multi_locate(self.h, k, C.byref(ctx), C.byref(slot), C.byref(dev)))
        return int(slot.value), int(dev.value)
