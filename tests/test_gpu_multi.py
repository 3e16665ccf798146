"""bf_multi_* (include/bf_cuda.h): the slice-sharded multi-GPU front -- block-cyclic deal of a batch,
one persistent launch per device, ONE NCCL all-gather of the per-slice flow records.
Results must be identical to the single-context batch, whatever the device count."""
import os
import subprocess

import numpy as np
import pytest

import better_flow_b200 as bf
from better_flow_b200 import synth
from helpers import same_model

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "better_flow_b200", "bf_motion_compensator")


def batch():
    st = synth.make_stream(240, 180, 2.0e6, 0.26, seed=23)
    return synth.cut_slices(st, 0.02)[:13]


def single(slices, ctx240):
    ctx240.reset()
    for s in slices:
        ctx240.add(s.fr_x, s.fr_y, s.t_ns, 3, 12)
    ctx240.run()
    return ctx240.results()


def same_results(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x["rc"] == y["rc"] and x["iters"] == y["iters"] and x["n_events"] == y["n_events"]
        assert same_model(x["model"], y["model"])      # (another grouping of CTAs: fp64 moment sums to the last bits)
        assert x["dividers"].tobytes() == y["dividers"].tobytes()


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
def test_multi_equals_single_context(ctx240, n_dev):
    if bf.load().bf_device_count() < n_dev:
        pytest.skip("needs %d CUDA devices" % n_dev)
    slices = batch()
    want = single(slices, ctx240)
    m = bf.MultiContext(n_dev, 180, 240, 3, max_events_per_device=1 << 20, max_slices_per_device=16)
    try:
        for rep in range(2):                      # the context is reusable
            m.reset()
            for s in slices:
                m.add_packed(bf.pack_events(s.fr_x, s.fr_y, s.t_ns), 3, 12)
            m.run()
            m.sync()
            same_results(m.results(), want)
        owners = [m.locate(k)[1] for k in range(len(slices))]
        assert owners == [(k // 4) % n_dev for k in range(len(slices))]
        assert m.launches == 2 * min(n_dev, (len(slices) + 3) // 4)
    finally:
        m.close()


def test_multi_large_sensor_needs_more_than_48k_shared_memory_on_every_device():
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per device: a 1280x720 context at scale 3 needs ~60 KB of
    dynamic shared memory, so every device of a bf_multi front must have had the attribute raised (round 1 raised
    it on device 0 only and the launch on devices 1.. failed with 'invalid argument')."""
    n_dev = min(2, bf.load().bf_device_count())
    st = synth.make_stream(1280, 720, 20e6, 0.02, seed=67)
    sls = synth.cut_slices(st, 0.0025)[:8]
    one = bf.Context(720, 1280, 3, max_events=len(st) + 16, max_slices=9, device=0)
    try:
        assert one.get_option("smem_bytes") > 48 * 1024
        for s in sls:
            one.add(s.fr_x, s.fr_y, s.t_ns, 3, 3)
        one.run()
        want = one.results()
    finally:
        one.close()
    m = bf.MultiContext(n_dev, 720, 1280, 3, max_events_per_device=len(st) + 16, max_slices_per_device=9)
    try:
        m.set_option("block", 1)                 # slices alternate between the devices
        for s in sls:
            m.add_packed(bf.pack_events(s.fr_x, s.fr_y, s.t_ns), 3, 3)
        m.run()
        m.sync()
        same_results(m.results(), want)
        if n_dev > 1:
            assert sorted(set(m.locate(k)[1] for k in range(len(sls)))) == [0, 1]
    finally:
        m.close()


def test_multi_ragged_batches(ctx240):
    """Fewer slices than devices x block, an empty batch, and a slice below the 1000-event guard."""
    n_dev = min(2, bf.load().bf_device_count())
    slices = batch()[:3]
    m = bf.MultiContext(n_dev, 180, 240, 3, max_events_per_device=1 << 20, max_slices_per_device=16)
    try:
        m.set_option("block", 1)
        m.reset(); m.run(); m.sync()
        assert m.size() == 0
        m.reset()
        for s in slices:
            m.add_packed(bf.pack_events(s.fr_x, s.fr_y, s.t_ns), 3, 12)
        tiny = slices[0]
        m.add_packed(bf.pack_events(tiny.fr_x[:500], tiny.fr_y[:500], tiny.t_ns[:500]), 3, 12)
        m.run(); m.sync()
        got = m.results()
        same_results(got[:3], single(slices, ctx240))
        assert got[3]["rc"] == bf.RC_SKIPPED and got[3]["n_events"] == 500
    finally:
        m.close()


def test_cli_gpus_flag_equals_single_gpu(tmp_path):
    if bf.load().bf_device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "better_flow_b200"), "cli"])
    st = synth.make_stream(240, 180, 2.0e6, 0.2, seed=29)
    rec = np.zeros(len(st), dtype=np.dtype([("t", "<u8"), ("x", "<u2"), ("y", "<u2"), ("p", "<u4")]))
    rec["t"], rec["x"], rec["y"], rec["p"] = st.t_ns, st.x, st.y, st.p
    binf = tmp_path / "s.bin"
    rec.tofile(binf)
    outs = []
    for gpus in (1, 2):
        out = tmp_path / ("f%d.txt" % gpus)
        r = subprocess.run([CLI, "--quiet", "--stm-disable", "--max-iter=10", "--batch=6", "--gpus=%d" % gpus,
                            "--flow-out=%s" % out, "-o", str(tmp_path / ("uv%d.txt" % gpus)), str(binf)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-1500:]
        outs.append((open(out).read(), open(tmp_path / ("uv%d.txt" % gpus)).read()))
    assert outs[0] == outs[1] and len(outs[0][0].splitlines()) >= 8


def test_tail_helping_is_result_neutral_and_engages():
    """When the slice queue runs dry, idle groups join slices that are still being minimised (the slice is
    then worked on by 2G CTAs).  Partial sums are combined in a different order, so results may differ in
    the last bits of the fp64 reductions but iteration counts, dividers and the flow (to 1e-12) must not."""
    st = synth.make_stream(240, 180, 3.0e6, 0.01 * 330, seed=31)
    sls = synth.cut_slices(st, 0.01)[:330]
    out = {}
    for help_on in (0, 1):
        ctx = bf.Context(180, 240, 3, max_events=len(st) + 64, max_slices=len(sls) + 1, device=0)
        try:
            ctx.set_option("tail_help", help_on)
            for s in sls:
                ctx.add(s.fr_x, s.fr_y, s.t_ns, 3, -1)
            ctx.run(want_events=True)
            res = ctx.results()
            ev = ctx.events(len(sls) - 1, len(sls[-1].fr_x))
            ctx.run()                                # and again: the images must have been left all-zero
            res2 = ctx.results()
            out[help_on] = (res, ev, res2)
        finally:
            ctx.close()
    a, b = out[0], out[1]
    for x, y, y2 in zip(a[0], b[0], b[2]):
        assert x["rc"] == y["rc"] == 0 and x["iters"] == y["iters"] == y2["iters"]
        assert x["dividers"].tobytes() == y["dividers"].tobytes()
        assert np.allclose(x["model"][7:11], y["model"][7:11], rtol=1e-12, atol=0)
        assert y["model"].tobytes() == y2["model"].tobytes() or np.allclose(y["model"], y2["model"], rtol=1e-12)
    assert np.allclose(a[1]["pr_x"], b[1]["pr_x"], rtol=0, atol=1e-9)


def test_tail_helping_with_big_grids_warm_starts_scales_and_event_readback(oracle_port):
    """Helping across everything a slice can carry: a 640x480 sensor (cell grid larger than one list
    chunk, so clearing goes through the flag re-scan), scales 1/3/5 in one batch, warm-started slices,
    per-event read-back -- and more groups than slices, so helpers join from the first iterations on.
    Every slice must equal its own single-slice run (where nobody can help) and the exact-sum oracle."""
    st = synth.make_stream(640, 480, 4.0e6, 0.05, seed=37)
    sls = synth.cut_slices(st, 0.005)[:10]
    ctx = bf.Context(480, 640, 5, max_events=len(st) + 64, max_slices=16, device=0)
    try:
        init = np.array([240.0, 320.0, 0, 0, 0, 0, 0, 0.01, -0.02, 1e-5, -2e-5])
        cfg = [((1, 3, 5)[k % 3], 6 + k, init if k % 4 == 3 else None) for k in range(len(sls))]
        singles = []
        for s, (scale, mi, ini) in zip(sls, cfg):
            singles.append(ctx.minimize(s.fr_x, s.fr_y, s.t_ns, scale, mi, init=ini, want_events=True))
        ctx.set_option("group_size", 2)          # 148 groups for 10 slices: 138 groups start as helpers
        ctx.reset()
        for s, (scale, mi, ini) in zip(sls, cfg):
            ctx.add(s.fr_x, s.fr_y, s.t_ns, scale, mi, init=ini)
        ctx.run(want_events=True)
        for k, (s, (scale, mi, ini), want) in enumerate(zip(sls, cfg, singles)):
            got = ctx.result(k)
            assert got["rc"] == want["rc"] == 0 and got["iters"] == want["iters"]
            assert np.allclose(got["model"][7:11], want["model"][7:11], rtol=1e-11, atol=0)
            ev = ctx.events(k, len(s.fr_x))
            assert np.allclose(ev["pr_x"], want["pr_x"], rtol=0, atol=1e-9) and np.allclose(ev["nx"], want["nx"], rtol=0, atol=1e-12)
            if k < 3:
                ex = oracle_port.minimize(s.fr_x, s.fr_y, s.t_ns, scale=scale, max_iter=mi, init_model=ini, rows=480, cols=640,
                                          accum_mode=1)
                assert ex["iters"] == got["iters"]
                assert np.allclose(got["model"][7:11], ex["model"][7:11], rtol=1e-9, atol=0)
        ctx.run()                                  # images were left all-zero: same answers again
        for k, want in enumerate(singles):
            assert ctx.result(k)["iters"] == want["iters"]
    finally:
        ctx.close()
