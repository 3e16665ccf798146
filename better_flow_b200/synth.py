"""Synthetic DAVIS-shaped event streams (workload generator for tests and bench.py).

The reference ships no datasets, so every workload is synthetic:
a set of random line segments ("contours") translating -- optionally also rotating and
expanding about the frame centre -- with sub-pixel jitter, sampled at random times.
This follows the generator spec of SURVEY.md section 8(d).

Coordinates follow the reference's *file* convention: ``x`` = column (0..cols-1),
``y`` = row (0..rows-1), exactly the ``t x y p`` text format that
``bf_motion_compensator`` reads (bf_motion_compensator.cpp:190-202).  The reference then
swaps them when it builds an ``Event(y, x, t)``: ``fr_x`` = row, ``fr_y`` = column.
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class Stream:
    """A time-ordered event stream in file convention (x = column, y = row)."""
    cols: int
    rows: int
    x: np.ndarray  # uint16, column
    y: np.ndarray  # uint16, row
    t_ns: np.ndarray  # int64, non-decreasing, first event near 0
    p: np.ndarray  # uint8 polarity

    def __len__(self) -> int:
        return int(self.t_ns.shape[0])

    def to_text(self, path: str, t_offset_s: float = 1.0) -> None:
        """Write the ``t x y p`` text format the reference CLI parses."""
        t = self.t_ns.astype(np.float64) * 1e-9 + t_offset_s
        with open(path, "w") as f:
            for i in range(len(self)):
                f.write("%.9f %d %d %d\n" % (t[i], self.x[i], self.y[i], self.p[i]))


def make_stream(cols: int, rows: int, rate_eps: float, duration_s: float, seed: int,
                vel=(80.0, -40.0), omega: float = 0.0, expand: float = 0.0) -> Stream:
    """Generate ``rate_eps * duration_s`` events.

    vel     (vx, vy) px/s in file coordinates (x = column, y = row)
    omega   rad/s rotation about the frame centre
    expand  1/s isotropic expansion about the frame centre
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    n = int(round(rate_eps * duration_s))
    n_seg = max(4, int(round(40.0 * (cols * rows) / (240.0 * 180.0))))
    ox = rng.uniform(0, cols, n_seg)
    oy = rng.uniform(0, rows, n_seg)
    ang = rng.uniform(0, np.pi, n_seg)
    length = rng.uniform(30.0, 60.0, n_seg)

    seg = rng.integers(0, n_seg, n)
    pos = rng.uniform(0.0, 1.0, n)
    t = np.sort(rng.uniform(0.0, duration_s, n))
    jx = rng.normal(0.0, 0.3, n)
    jy = rng.normal(0.0, 0.3, n)
    pol = (rng.uniform(0.0, 1.0, n) < 0.5).astype(np.uint8)

    px = ox[seg] + np.cos(ang[seg]) * length[seg] * pos + jx
    py = oy[seg] + np.sin(ang[seg]) * length[seg] * pos + jy
    if omega != 0.0 or expand != 0.0:
        cx, cy = cols / 2.0, rows / 2.0
        rx, ry = px - cx, py - cy
        a = omega * t
        k = np.exp(expand * t)
        px = cx + k * (np.cos(a) * rx - np.sin(a) * ry)
        py = cy + k * (np.sin(a) * rx + np.cos(a) * ry)
    px = px + vel[0] * t
    py = py + vel[1] * t
    xi = np.floor(np.mod(px, cols)).astype(np.int64)
    yi = np.floor(np.mod(py, rows)).astype(np.int64)
    xi = np.clip(xi, 0, cols - 1).astype(np.uint16)
    yi = np.clip(yi, 0, rows - 1).astype(np.uint16)
    t_ns = np.floor(t * 1e9).astype(np.int64)
    return Stream(cols, rows, xi, yi, t_ns, pol)


@dataclasses.dataclass
class Slice:
    """One independent time slice in the reference's *event* convention.

    fr_x = row, fr_y = column (Event(y, x, t), bf_motion_compensator.cpp:200); events are in the
    order OptimizerRolling iterates them, newest -> oldest (dvs_flow.h:196-198,
    datastructures.h:86-96); ``t_ns`` is the local time ``timestamp - slice_start``
    (event.h:61-63).
    """
    fr_x: np.ndarray  # uint16
    fr_y: np.ndarray  # uint16
    t_ns: np.ndarray  # int32
    rows: int
    cols: int


def cut_slices(stream: Stream, slice_s: float, max_events: int | None = None,
               min_events: int = 1) -> list[Slice]:
    """Cut a stream into consecutive non-overlapping slices (stm-disabled semantics)."""
    span = int(round(slice_s * 1e9))
    out = []
    t = stream.t_ns
    k = 0
    t0 = 0
    while k < len(stream):
        hi = int(np.searchsorted(t, t0 + span, side="left"))
        lo = k
        if max_events is not None and hi - lo > max_events:
            lo = hi - max_events
        if hi - lo >= min_events:
            sl = slice(lo, hi)
            order = np.arange(hi - 1, lo - 1, -1)
            out.append(Slice(
                fr_x=np.ascontiguousarray(stream.y[order]),
                fr_y=np.ascontiguousarray(stream.x[order]),
                t_ns=np.ascontiguousarray((t[order] - t0).astype(np.int32)),
                rows=stream.rows, cols=stream.cols))
            del sl
        k = hi
        t0 += span
    return out
