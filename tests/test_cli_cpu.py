"""bf_motion_compensator without a GPU: the flag surface of the reference's tool (bf_motion_compensator.cpp:64-130)
plus the additions SURVEY 8b asks for, and the no-CPU-fallback rule (the tool must fail loudly, not compute on the host)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "better_flow_b200", "bf_motion_compensator")

REFERENCE_FLAGS = ["--refresh-time=", "--refresh-event-count=", "-i/--interactive", "-G", "--stm-disable", "--img", "--img-prefix",
                   "--video", "--video-name", "--video-fps=", "--bufferize-file", "-o <name>/--outfile=", "--version"]
ADDED_FLAGS = ["--max-iter=", "--scale=", "--slice-time=", "--max-events=", "--sensor=", "--flow-out=", "--batch=", "--gpus=",
               "--optimizer=", "--device=", "--quiet"]


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "better_flow_b200"), "all"])
    return CLI


def test_help_lists_the_reference_flags_and_the_additions(cli):
    r = subprocess.run([cli, "--help"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0
    text = r.stdout + r.stderr
    for f in REFERENCE_FLAGS + ADDED_FLAGS:
        assert f in text, f
    # the reference's defaults (bf_motion_compensator.cpp:6-7,44-45): 50000 events, 200 ms, 33 ms, 20000 events
    assert "50000 events" in text and "0.200000 seconds" in text
    assert "default = 0.033000" in text and "default = 20000" in text


def test_version_and_missing_file(cli):
    r = subprocess.run([cli, "--version"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "better flow" in r.stdout
    r = subprocess.run([cli], capture_output=True, text=True, timeout=60)
    assert "usage" in (r.stdout + r.stderr)


def test_no_cpu_fallback(cli, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    f = tmp_path / "ev.txt"
    with open(f, "w") as o:
        for k in range(3000):
            o.write("%.9f %d %d %d\n" % (1.0 + k * 1e-5, k % 240, (7 * k) % 180, k & 1))
    r = subprocess.run([cli, "--quiet", "--refresh-event-count=1000", str(f)], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)


# ---- the whole tool against the reference's tool, on the CPU ----------------------------------------------------
# bf_motion_compensator.cpp is linked a second time against the oracle-backed test double of the C ABI
# (tests/cpu/mock_bf_cuda.cpp) instead of the CUDA library.  With an exact back end, everything the tool itself does --
# parsing, triggers, ring buffer, warm start, per-event flow, accumulation and de-duplication, the -o writer, the
# per-slice dumps -- must reproduce the reference tool's output byte for byte.

REF_CLI = os.path.join(ROOT, "oracle", "_ref", "bf_motion_compensator_ref")


@pytest.fixture(scope="module")
def mock_cli():
    exe = os.path.join(ROOT, "tests", "cpu", "bf_motion_compensator_mock")
    inc = os.path.join(ROOT, "better_flow_b200", "include")
    odir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["make", "-s", "-C", odir, "port"])
    srcs = [os.path.join(ROOT, "better_flow_b200", "src", "bf_motion_compensator.cpp"), os.path.join(ROOT, "tests", "cpu", "mock_bf_cuda.cpp")]
    deps = srcs + [os.path.join(inc, "better_flow", f) for f in os.listdir(os.path.join(inc, "better_flow"))] + \
        [os.path.join(ROOT, "include", "bf_cuda.h"), os.path.join(odir, "libbf_oracle.so")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(f) for f in deps):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-pthread", "-I" + inc, "-I" + os.path.join(ROOT, "include"),
                               *srcs, "-L" + odir, "-lbf_oracle", "-Wl,-rpath," + odir, "-o", exe])
    return exe


def _stable(stdout, *paths):
    """stdout without the wall-clock lines and with the per-run file names masked."""
    out = []
    for l in stdout.splitlines():
        if "Elapsed:" in l or " sec\t" in l:
            continue
        for p in paths:
            l = l.replace(str(p), "<file>")
        out.append(l)
    return out


@pytest.mark.parametrize("extra", [[], ["--stm-disable"], ["--refresh-event-count=7000", "--refresh-time=0.011"]])
def test_tool_with_exact_back_end_equals_the_reference_tool_byte_for_byte(mock_cli, tmp_path, extra):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/bf_motion_compensator_ref not built (no /root/reference here)")
    from better_flow_b200 import synth
    st = synth.make_stream(240, 180, 0.7e6, 0.05, seed=17, vel=(90.0, -50.0), omega=0.4)
    txt = tmp_path / "events.txt"
    st.to_text(str(txt))
    o_ref, o_new = tmp_path / "ref_uv.txt", tmp_path / "new_uv.txt"
    ref = subprocess.run([REF_CLI] + extra + ["-o", str(o_ref), str(txt)], capture_output=True, text=True, timeout=600)
    new = subprocess.run([mock_cli] + extra + ["-o", str(o_new), str(txt)], capture_output=True, text=True, timeout=600)
    assert ref.returncode == 0 and new.returncode == 0, (ref.stderr[-500:], new.stderr[-1500:])
    assert open(o_ref, "rb").read() == open(o_new, "rb").read()          # t x y 1 v u for every accumulated event
    assert os.path.getsize(o_ref) > 100000
    a, b = _stable(ref.stdout, o_ref, txt), _stable(new.stdout, o_new, txt)
    assert len(a) > 20 and a == b                                         # per-slice model dumps, accounting lines
    # without -o the tool skips the per-event read-back (set_lazy_events): the dumps must not change
    lazy = subprocess.run([mock_cli] + extra + [str(txt)], capture_output=True, text=True, timeout=600)
    assert lazy.returncode == 0
    dumps = [l for l in b if l.startswith(("C:", "\t Shift", "\t Rot", "\t Div", "\t cnt"))]
    assert dumps and dumps == [l for l in _stable(lazy.stdout, txt) if l.startswith(("C:", "\t Shift", "\t Rot", "\t Div", "\t cnt"))]


def test_batched_tool_writes_the_same_flow_file(mock_cli, tmp_path):
    """--stm-disable --batch=N defers the minimisation of independent slices; the per-event flow written with -o
    (mapped back from the newest-first slice snapshots) and the per-slice flow lines are those of the unbatched run."""
    from better_flow_b200 import synth
    st = synth.make_stream(240, 180, 0.8e6, 0.08, seed=23)
    txt = tmp_path / "events.txt"
    st.to_text(str(txt))
    outs = []
    for batch in (1, 3):
        uv, fl = tmp_path / ("uv%d.txt" % batch), tmp_path / ("flow%d.txt" % batch)
        r = subprocess.run([mock_cli, "--quiet", "--stm-disable", "--max-iter=12", "--refresh-event-count=9000", "--batch=%d" % batch,
                            "--flow-out=%s" % fl, "-o", str(uv), str(txt)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-1500:]
        outs.append((open(uv, "rb").read(), open(fl).read()))
    assert outs[0] == outs[1] and len(outs[0][1].splitlines()) >= 6 and len(outs[0][0]) > 100000


def test_out_of_order_input_still_equals_the_reference_tool(mock_cli, tmp_path):
    """Timestamps that DEcrease here and there (sensor jitter, merged recordings): the -o aggregation falls back to
    the reference's literal nested scan (dvs_flow.h:350-389), whose result the indexed scan only reproduces for sorted
    input -- the files must stay byte-identical."""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/bf_motion_compensator_ref not built (no /root/reference here)")
    import numpy as np
    from better_flow_b200 import synth
    st = synth.make_stream(240, 180, 0.6e6, 0.05, seed=29, vel=(-70.0, 30.0))
    t = st.t_ns.copy()
    rng = np.random.default_rng(4)
    for k in rng.integers(100, len(t) - 100, 60):          # swap 60 neighbouring pairs and push a few events 2 ms back
        t[k], t[k + 1] = t[k + 1], t[k]
    for k in rng.integers(2000, len(t) - 100, 12):
        t[k] = max(0, t[k] - 2_000_000)
    st.t_ns = t
    txt = tmp_path / "events.txt"
    st.to_text(str(txt))
    o_ref, o_new = tmp_path / "ref_uv.txt", tmp_path / "new_uv.txt"
    ref = subprocess.run([REF_CLI, "-o", str(o_ref), str(txt)], capture_output=True, text=True, timeout=600)
    new = subprocess.run([mock_cli, "-o", str(o_new), str(txt)], capture_output=True, text=True, timeout=600)
    assert ref.returncode == 0 and new.returncode == 0, (ref.stderr[-500:], new.stderr[-1500:])
    assert open(o_ref, "rb").read() == open(o_new, "rb").read() and os.path.getsize(o_ref) > 100000
    assert _stable(ref.stdout, o_ref, txt) == _stable(new.stdout, o_new, txt)


def test_aggregation_of_bursty_pixels_over_many_buffers_equals_the_reference_tool(mock_cli, tmp_path):
    """-o keeps, of the copies the overlapping buffers hold, the first one and drops from LATER buffers every copy of the
    same pixel that is not newer and closer than 0.1 ms (dvs_flow.h:350-389, event.h:40-45).  Hot pixels that fire in
    bursts (gaps on both sides of 0.1 ms, exactly 0.1 ms, equal timestamps) over a dozen overlapping buffers: the
    indexed scan here must drop exactly what the reference's nested scan drops (which copy of an event survives decides
    which slice's flow the file reports for it) -- byte-identical files."""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/bf_motion_compensator_ref not built (no /root/reference here)")
    import numpy as np
    from better_flow_b200 import synth
    st = synth.make_stream(240, 180, 0.4e6, 0.04, seed=31, vel=(60.0, 20.0))
    rng = np.random.default_rng(8)
    hot = [(int(rng.integers(20, 220)), int(rng.integers(20, 160))) for _ in range(24)]
    bx, by, bt = [], [], []
    for _ in range(1300):
        x, y = hot[int(rng.integers(0, len(hot)))]
        t0 = int(rng.integers(0, 39_000_000))
        gaps = rng.choice([0, 1, 40_000, 99_999, 100_000, 100_001, 150_000, 30_000], size=3)
        t = t0
        for g in [0] + list(gaps):
            t += int(g)
            bx.append(x); by.append(y); bt.append(t)
    t_all = np.concatenate([st.t_ns, np.array(bt, dtype=np.int64)])
    order = np.argsort(t_all, kind="stable")
    st.x = np.concatenate([st.x, np.array(bx, dtype=st.x.dtype)])[order]
    st.y = np.concatenate([st.y, np.array(by, dtype=st.y.dtype)])[order]
    st.p = np.concatenate([st.p, np.ones(len(bt), dtype=st.p.dtype)])[order]
    st.t_ns = t_all[order]
    txt = tmp_path / "events.txt"
    st.to_text(str(txt))
    o_ref, o_new = tmp_path / "ref_uv.txt", tmp_path / "new_uv.txt"
    flags = ["--refresh-event-count=1800"]
    ref = subprocess.run([REF_CLI] + flags + ["-o", str(o_ref), str(txt)], capture_output=True, text=True, timeout=900)
    new = subprocess.run([mock_cli] + flags + ["-o", str(o_new), str(txt)], capture_output=True, text=True, timeout=900)
    assert ref.returncode == 0 and new.returncode == 0, (ref.stderr[-500:], new.stderr[-1500:])
    n_buffers = sum(1 for l in new.stdout.splitlines() if l.startswith("\tBuffer: "))
    kept = len(open(o_new, "rb").read().splitlines())
    assert n_buffers >= 12 and kept == len(st)                    # (every event is emitted once, from the first buffer that holds it)
    assert open(o_ref, "rb").read() == open(o_new, "rb").read()
    assert _stable(ref.stdout, o_ref, txt) == _stable(new.stdout, o_new, txt)
