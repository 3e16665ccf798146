"""Text event reader: better_flow/event_file.h's block reader must yield exactly what the
reference's `ifstream >> t >> x >> y >> p` loop yields (bf_motion_compensator.cpp:190-202) -- same
doubles, same stopping point -- only faster."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from better_flow_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def rd():
    so = os.path.join(HERE, "cpu", "libreader_shim.so")
    src = os.path.join(HERE, "cpu", "reader_shim.cpp")
    inc = os.path.join(ROOT, "better_flow_b200", "include")
    hdrs = [os.path.join(inc, "better_flow", f) for f in os.listdir(os.path.join(inc, "better_flow"))]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-I" + inc,
                               "-I" + os.path.join(ROOT, "include"), src, "-o", so])
    lib = C.CDLL(so)
    for f in (lib.rd_fast, lib.rd_iostream, lib.rd_from_file, lib.rd_prefetch):
        f.restype = C.c_longlong
    lib.wr_fast.restype = lib.wr_iostream.restype = C.c_double
    return lib


def parse(lib, fn, path, cap):
    t = np.zeros(cap, np.float64); x = np.zeros(cap, np.uint32); y = np.zeros(cap, np.uint32); p = np.zeros(cap, np.uint8)
    secs = C.c_double()
    n = fn(str(path).encode(), C.c_longlong(cap), t.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p),
           y.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), C.byref(secs))
    return n, t[:n], x[:n], y[:n], p[:n], secs.value


def same(a, b):
    assert a[0] == b[0]
    assert a[1].tobytes() == b[1].tobytes()      # bit-identical doubles
    for u, v in zip(a[2:5], b[2:5]):
        assert np.array_equal(u, v)


def test_reader_equals_iostream_on_a_synthetic_stream(rd, tmp_path):
    st = synth.make_stream(240, 180, 3e6, 0.1, seed=11)
    path = tmp_path / "ev.txt"
    st.to_text(str(path))
    cap = len(st) + 10
    a = parse(rd, rd.rd_fast, path, cap)
    b = parse(rd, rd.rd_iostream, path, cap)
    assert a[0] == len(st)
    same(a, b)
    same(parse(rd, rd.rd_prefetch, path, cap), b)            # background-thread variant: same records, same order
    assert parse(rd, rd.rd_prefetch, path, 1000)[0] == 1000   # the consumer may stop early (the worker is joined cleanly)
    print("parse %d events: block reader %.3f s, iostream %.3f s (%.1fx)" % (a[0], a[5], b[5], b[5] / a[5]))
    assert a[5] < b[5]


@pytest.mark.parametrize("text,expect", [
    ("", 0),
    ("\n\n", 0),
    ("1.5 3 4 1", 1),                                             # no trailing newline
    ("1.5 3 4 1\n2.5 5 6 0\n", 2),
    ("  1.5\t3   4 1 \r\n2.5 5 6 0", 2),                          # mixed whitespace, CRLF
    ("1.5 3\n4 1 2.5\n5 6 0\n", 2),                               # records spanning lines (operator>> does not care)
    ("1e-3 3 4 1\n+2.5E0 +5 6 0\n.5 1 2 1\n7. 1 2 0\n", 4),        # exponent forms, leading '+', bare dot forms
    ("1.5 3 4 1\n2.5 5 6 2\n3.5 1 1 1\n", 1),                     # p = 2 is not a bool: stop there
    ("1.5 3 4 1\nabc 5 6 0\n3.5 1 1 1\n", 1),                     # garbage: stop there
    ("1.5 3 4 1\n2.5 -5 6 0\n", 2),                               # negative coordinate wraps like strtoul
    ("1.5 3 4 1\n2.5 5 6\n", 1),                                  # truncated last record
    ("0.000000001 0 0 0\n1234567890.123456789 239 179 1\n", 2),
    ("1.5 3 4 1\n2.5 99999999999 6 0\n", 1),                      # coordinate overflows uint
])
def test_reader_edge_cases_equal_iostream(rd, tmp_path, text, expect):
    path = tmp_path / "e.txt"
    path.write_text(text)
    a = parse(rd, rd.rd_fast, path, 16)
    b = parse(rd, rd.rd_iostream, path, 16)
    assert b[0] == expect, "iostream baseline changed its mind"
    same(a, b)
    same(parse(rd, rd.rd_prefetch, path, 16), b)


def test_reader_long_lines_and_block_boundaries(rd, tmp_path):
    # records separated by long runs of blanks so that tokens straddle the 4 MiB block boundary
    rng = np.random.default_rng(5)
    parts = []
    for i in range(3000):
        parts.append("%.9f%s%d %d %d%s" % (1.0 + i * 1e-4, " " * int(rng.integers(1, 4000)), rng.integers(0, 240),
                                           rng.integers(0, 180), rng.integers(0, 2), "\n" if i % 7 else " "))
    path = tmp_path / "long.txt"
    path.write_text("".join(parts))
    a = parse(rd, rd.rd_fast, path, 4000)
    b = parse(rd, rd.rd_iostream, path, 4000)
    assert b[0] == 3000
    same(a, b)


def test_missing_file_yields_nothing(rd, tmp_path):
    a = parse(rd, rd.rd_fast, tmp_path / "nope.txt", 4)
    assert a[0] == 0


def test_from_file_rebases_time_and_swaps_xy(rd, tmp_path):
    path = tmp_path / "e.txt"
    path.write_text("10.5 7 3 1\n10.500001 8 4 0\n10.75 9 5 1\n")
    ts = np.zeros(8, np.uint64); fx = np.zeros(8, np.uint32); fy = np.zeros(8, np.uint32)
    n = rd.rd_from_file(str(path).encode(), C.c_longlong(8), ts.ctypes.data_as(C.c_void_p), fx.ctypes.data_as(C.c_void_p),
                        fy.ctypes.data_as(C.c_void_p))
    assert n == 3
    # Event(row = file y, column = file x, FROM_SEC(t - t0)) with FROM_SEC(x) = ull(1e9 * x)
    want = [0, int(1000000000 * (10.500001 - 10.5)), int(1000000000 * (10.75 - 10.5))]
    assert ts[:3].tolist() == want
    assert fx[:3].tolist() == [3, 4, 5] and fy[:3].tolist() == [7, 8, 9]


def test_fast_record_path_never_diverges_from_iostream(rd, tmp_path):
    """Randomised record shapes around the fast path's acceptance rules (digit counts across the 15-digit limit,
    bare-dot forms, exponents, signs, tabs, CRLF, blank lines, several blanks, leading zeros, 9/10-digit coordinates):
    count and every value must equal what `ifstream >>` reads."""
    rng = np.random.default_rng(20261017)

    def ts():
        k = rng.integers(0, 8)
        ip = str(rng.integers(0, 10 ** int(rng.integers(1, 12))))
        fr = "".join(str(d) for d in rng.integers(0, 10, int(rng.integers(0, 12))))
        if k == 0: return ip                                  # integer only
        if k == 1: return ip + "."                            # bare trailing dot
        if k == 2: return "." + (fr or "5")                   # bare leading dot
        if k == 3: return "%s.%se%d" % (ip, fr or "0", rng.integers(-5, 6))
        if k == 4: return "+" + ip + "." + fr
        if k == 5: return "000" + ip + "." + fr
        return ip + "." + fr

    def coord():
        k = rng.integers(0, 10)
        if k == 0: return str(rng.integers(10 ** 8, 4294967295))   # 9-10 digits
        if k == 1: return "+" + str(rng.integers(0, 1000))
        if k == 2: return "00" + str(rng.integers(0, 1000))
        return str(rng.integers(0, 1280))

    seps = [" ", " ", " ", "  ", "\t", " \t "]
    eols = ["\n", "\n", "\n", "\r\n", " \n", "\n\n", "\t\n"]
    lines = []
    for _ in range(20000):
        s = lambda: seps[rng.integers(0, len(seps))]
        lines.append(ts() + s() + coord() + s() + coord() + s() + str(rng.integers(0, 2)) + eols[rng.integers(0, len(eols))])
    path = tmp_path / "fuzz.txt"
    path.write_text("".join(lines))
    a = parse(rd, rd.rd_fast, path, 30000)
    b = parse(rd, rd.rd_iostream, path, 30000)
    assert b[0] == 20000
    same(a, b)
    same(parse(rd, rd.rd_prefetch, path, 30000), b)


def test_flow_writer_equals_ostream_byte_for_byte(rd, tmp_path):
    """EventFile::to_file_uv (the -o file: "t x y 1 v u", 9 decimals) formats with std::to_chars into 1 MB blocks; the
    file must be what the reference's `ofstream << fixed << setprecision(9)` writes (event_file.h:265-289), including
    -0, nan, inf, huge and tiny values, and lines that fall on a block boundary."""
    rng = np.random.default_rng(5)
    n = 120_000                                                   # > 4 blocks of 1 MB
    ts = np.sort(rng.integers(0, 2 * 10 ** 12, n)).astype(np.uint64)
    fx = rng.integers(0, 720, n).astype(np.uint32); fy = rng.integers(0, 1280, n).astype(np.uint32)
    bu = rng.normal(0, 300, n); bv = rng.normal(0, 1e-4, n)
    bu[::97] = 0.0; bv[::89] = -0.0; bu[5] = np.nan; bv[6] = np.inf; bu[7] = -np.inf; bv[8] = 1e300; bu[9] = 5e-10; bv[10] = -5e-10
    bu[11] = 123456789012345678.0; bv[12] = 0.9999999995; bu[13] = 2.5e-9; bv[14] = 0.1234567885
    bits = rng.integers(0, 2 ** 63, 2000).astype(np.uint64).view(np.float64)
    bu[1000:3000] = np.where(np.isfinite(bits), bits, 1.0)        # arbitrary bit patterns (incl. subnormals, 1e+-300)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    a, b = tmp_path / "fast.txt", tmp_path / "ios.txt"
    s_fast = rd.wr_fast(str(a).encode(), C.c_longlong(n), p(ts), p(fx), p(fy), p(bu), p(bv))
    s_ios = rd.wr_iostream(str(b).encode(), C.c_longlong(n), p(ts), p(fx), p(fy), p(bu), p(bv))
    da, db = open(a, "rb").read(), open(b, "rb").read()
    assert len(da) > 4 * (1 << 20) and da == db
    assert da.splitlines()[0].split() == [("%.9f" % (ts[0] / 1e9)).encode(), str(fy[0]).encode(), str(fx[0]).encode(), b"1",
                                          ("%.9f" % bv[0]).encode(), ("%.9f" % bu[0]).encode()]
    print("to_file_uv %.3f s, ostream %.3f s for %d events" % (s_fast, s_ios, n))
