"""Host-side unit tests of better_flow_b200/csrc/bf_logic.h -- the geometry, packed-accumulator and
gradient-descent control logic that the kernel executes -- compiled for the CPU and compared with
the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import golden, case_events

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lg():
    so = os.path.join(HERE, "cpu", "liblogic_shim.so")
    src = os.path.join(HERE, "cpu", "logic_shim.cpp")
    hdr = os.path.join(HERE, "..", "better_flow_b200", "csrc", "bf_logic.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", so])
    lib = C.CDLL(so)
    lib.lg_accumulate.restype = C.c_float
    return lib


def test_geometry_matches_set_cloud(lg, oracle_port):
    G, EV = golden()
    rng = np.random.default_rng(3)
    boxes = [(0, 179, 0, 239), (12, 101, 30, 200), (5, 6, 7, 9), (0, 0, 0, 0), (17, 170, 3, 238)]
    boxes += [tuple(sorted(rng.integers(0, 180, 2)) + sorted(rng.integers(0, 240, 2))) for _ in range(20)]
    for (x0, x1, y0, y1) in boxes:
        for scale in (1, 3, 5, 7):
            fx = np.array([x0, x1], dtype=np.uint16)
            fy = np.array([y0, y1], dtype=np.uint16)
            su = oracle_port.setup_slice(fx, fy, 180, 240, scale)
            ints = (C.c_int * 6)()
            dbls = (C.c_double * 2)()
            lg.lg_geom(int(x0), int(x1), int(y0), int(y1), scale, ints, dbls)
            assert (ints[0], ints[1], ints[2], ints[3]) == (su.wsize_x, su.wsize_y, su.img_rows, su.img_cols)
            assert (dbls[0], dbls[1]) == (su.x_shift, su.y_shift)
            assert (ints[4], ints[5]) == (int(su.x_shift), int(su.y_shift))


def test_full_frame_davis240_shift_is_2p5(lg):
    ints = (C.c_int * 6)()
    dbls = (C.c_double * 2)()
    lg.lg_geom(0, 179, 0, 239, 3, ints, dbls)
    assert (ints[2], ints[3]) == (540, 720) and (dbls[0], dbls[1]) == (2.5, 2.5) and (ints[4], ints[5]) == (2, 2)


def test_tiny_window_guard(lg):
    # optimizer_rolling.h:49 -- both dimensions below scale*RES/15
    assert lg.lg_guard_tiny(10, 17, 20, 30, 3, 180, 240) == 1
    assert lg.lg_guard_tiny(10, 17, 20, 120, 3, 180, 240) == 0
    assert lg.lg_guard_tiny(0, 179, 0, 239, 3, 180, 240) == 0


@pytest.mark.parametrize("n,span", [(30000, 10_000_000), (90000, 30_000_000), (100000, 50_000_000),
                                    (200000, 20_000_000), (1000000, 10_000_000), (50000, 200_000_000)])
def test_packed_accumulator_is_exact_at_baseline_sizes(lg, n, span):
    """No BASELINE configuration needs time quantisation: count and sum of t fit in 64 bits even if
    every event of the slice landed on one pixel."""
    cs, q = C.c_int(), C.c_int()
    lg.lg_pack(n, 0, span - 1, C.byref(cs), C.byref(q))
    assert q.value == 0
    cnt_bits = 64 - cs.value
    assert (1 << cnt_bits) > n
    assert n * (span - 1) < (1 << cs.value)


def test_packed_accumulator_quantises_only_beyond_64_bits(lg):
    cs, q = C.c_int(), C.c_int()
    lg.lg_pack(1 << 22, 0, (1 << 31) - 1, C.byref(cs), C.byref(q))   # 31 + 2*23 = 77 bits -> q = 13
    assert q.value == 13


def test_pack_offset_and_fast_path_selection(lg):
    """Non-negative local times are packed without an offset (and take the device's fast unpack when
    every possible box sum stays below 2^52); negative times or over-long spans re-base to t_min."""
    def pack(n, lo, hi):
        o = (C.c_int * 4)()
        lg.lg_pack_full(n, lo, hi, o)
        return tuple(o)
    for n, span in [(30000, 10_000_000), (90000, 30_000_000), (100000, 50_000_000), (200000, 20_000_000),
                    (1000000, 10_000_000), (50000, 200_000_000)]:
        cs, q, off, fast = pack(n, 1234, span)
        assert (q, off, fast) == (0, 0, 1)
        assert n * span < (1 << 52)
    assert pack(90000, -5, 30_000_000)[2:] == (-5, 0)                 # negative time: offset, generic path
    cs, q, off, fast = pack(1 << 16, 1_000_000_000, 2_000_000_000)    # 31 + 2*17 > 64 as is, fits after re-basing
    assert (q, off, fast) == (0, 1_000_000_000, 0)
    cs, q, off, fast = pack(1 << 14, 0, (1 << 31) - 1)                # 31 + 15 = 46 bits: fast
    assert (q, off, fast) == (0, 0, 1)
    cs, q, off, fast = pack((1 << 21) - 1, 0, (1 << 22) - 1)          # fits in 64 but sums may reach 2^43 < 2^52: fast
    assert (q, off, fast) == (0, 0, 1)


def test_unpack_equals_exact_mean(lg):
    rng = np.random.default_rng(1)
    for _ in range(200):
        k = int(rng.integers(1, 60))
        t = rng.integers(-2_000_000, 30_000_000, k).astype(np.int32)
        got = lg.lg_accumulate(90000, -2_000_000, 30_000_000, k, t.ctypes.data_as(C.POINTER(C.c_int)))
        s = np.float32(np.float64(int(t.astype(np.int64).sum())) / 1e9)
        want = np.float32(s / np.float32(k))
        assert np.float32(got) == want
        # the same events with non-negative times (no offset in the packed word)
        t2 = (t.astype(np.int64) + 2_000_000).astype(np.int32)
        got2 = lg.lg_accumulate(90000, 0, 32_000_000, k, t2.ctypes.data_as(C.POINTER(C.c_int)))
        s2 = np.float32(np.float64(int(t2.astype(np.int64).sum())) / 1e9)
        assert np.float32(got2) == np.float32(s2 / np.float32(k))


def test_gd_control_flow_replay_matches_oracle(lg, oracle_port):
    """Record the per-iteration image sums of an oracle run (exact-sum mode) and replay them through
    bf_opt_advance: iteration count, dividers and accumulated model must come out identical."""
    G, EV = golden()
    case = next(c for c in G["cases"] if c["name"] == "davis240_10ms_converge")
    fx, fy, t, noise, init = case_events(case)
    scale = case["scale"]
    su = oracle_port.setup_slice(fx, fy, 180, 240, scale)
    rows, cols = su.img_rows, su.img_cols
    i0, j0 = rows // 2, cols // 2
    # drive the oracle one iteration at a time through its public pieces
    pr_x, pr_y = fx.astype(np.float64), fy.astype(np.float64)
    tl = t.astype(np.int64)
    sums = []
    model = np.zeros(11)
    xd = yd = np.float32(1.0); rd = dd = np.float32(10000.0)
    want = oracle_port.minimize(fx, fy, t, scale=scale, max_iter=-1, accum_mode=1)
    for it in range(want["iters"]):
        img = oracle_port.time_img(pr_x, pr_y, tl, su.wsize_x, su.wsize_y, scale, int(su.x_shift), int(su.y_shift), accum_mode=1)
        m7, gx, gy = oracle_port.model(img, want_grad=True)
        occ = img > np.float32(1e-6)
        ii, jj = np.nonzero(occ)
        gxo, gyo = gx[occ].astype(np.float64), gy[occ].astype(np.float64)
        sums.append([occ.sum(), ii.sum(), jj.sum(), gxo.sum(), gyo.sum(), ((ii - i0) * gxo).sum(), ((jj - j0) * gxo).sum(),
                     ((ii - i0) * gyo).sum(), ((jj - j0) * gyo).sum()])
        # advance the oracle's state exactly like iteration_step does, using its own projection
        out = np.zeros(17)
        arr = np.ascontiguousarray(np.array(sums, dtype=np.float64))
        k = lg.lg_replay(len(sums), arr.ctypes.data_as(C.POINTER(C.c_double)), su.x_min, su.x_max, su.y_min, su.y_max,
                         scale, i0, j0, -1, 100000, out.ctypes.data_as(C.POINTER(C.c_double)))
        pr_x, pr_y, _, _ = oracle_port.project(fx, fy, tl, pr_x, pr_y, -out[7], -out[8], out[0], out[1], out[10], -out[9])
    assert k == want["iters"] and int(out[16]) == want["iters"]
    assert np.array_equal(out[11:15].astype(np.float32), want["dividers"])
    assert np.allclose(out[7:11], want["model"][7:11], rtol=1e-9, atol=0)
    assert out[6] == want["model"][6]


def test_packed_events_are_checked_against_the_sensor(lg):
    """bf_batch_add_packed / bf_batch_add_staged refuse coordinates beyond the context's sensor (the device
    sizes the images from the events' bounding box); the noise bit of fr_y is not a coordinate bit."""
    import better_flow_b200 as bf
    rng = np.random.default_rng(3)
    n = 4097
    fx = rng.integers(0, 180, n).astype(np.uint16)
    fy = rng.integers(0, 240, n).astype(np.uint16)
    t = rng.integers(0, 30_000_000, n).astype(np.int32)
    noise = (rng.random(n) < 0.1).astype(np.uint8)
    ev = np.ascontiguousarray(bf.pack_events(fx, fy, t, noise))
    assert ev.dtype.itemsize == 8

    def ok(a, rx=180, ry=240):
        return lg.lg_events_in_sensor(a.ctypes.data_as(C.c_void_p), C.c_longlong(len(a)), rx, ry)

    assert ok(ev) == 1
    assert ok(ev[:0]) == 1
    assert ok(ev, 179, 240) == (1 if fx.max() < 179 else 0)
    for pos in (0, n // 2, n - 1):
        bad = ev.copy()
        raw = bad.view(np.uint16).reshape(-1, 4)
        raw[pos, 0] = 180                       # fr_x == res_x
        assert ok(bad) == 0
        bad = ev.copy()
        raw = bad.view(np.uint16).reshape(-1, 4)
        raw[pos, 1] = 240 | 0x8000              # fr_y == res_y, noise bit set
        assert ok(bad) == 0
        bad = ev.copy()
        raw = bad.view(np.uint16).reshape(-1, 4)
        raw[pos, 1] = 239 | 0x8000              # in range + noise bit: fine
        assert ok(bad) == 1
