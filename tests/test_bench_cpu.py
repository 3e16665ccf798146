"""bench.py on a box without a GPU: the reference arm (the reference's CPU path on the host cores), the
configuration table, and the no-CPU-fallback rule of the product arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run_bench(*args, env=None):
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, env=e, timeout=600)


def test_config_table_matches_baseline_json():
    sys.path.insert(0, ROOT)
    import bench
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert bench.METRIC.split()[0] in base["metric"]
    # (cols, rows, slice length) of BASELINE.json configs[1..4]
    want = {"cfg2": (240, 180, 0.030), "cfg3": (346, 260, 0.050), "cfg4": (640, 480, 0.020), "cfg5": (1280, 720, 0.010)}
    for name, (cols, rows, slice_s) in want.items():
        bench.select_config(name)
        assert (bench.SENSOR_COLS, bench.SENSOR_ROWS, bench.SLICE_S) == (cols, rows, slice_s)
        assert "%dx%d" % (cols, rows) in bench.WORKLOAD
        assert bench.MAX_ITER == -1 and bench.SCALE == 3                      # GD to convergence, reference default scale
        assert bench.RATE_EPS * bench.SLICE_S * bench.SLICES_PER_STEP * 8 > 50e6   # tens of MB of events per step
    bench.select_config("cfg2", slices=7)
    assert bench.SLICES_PER_STEP == 7 and "7 slices per step" in bench.WORKLOAD
    bench.select_config("cfg2")


def test_reference_arm_prints_one_json_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mevents/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("DAVIS-240C")


def test_reference_arm_other_ranks_do_no_work():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2",
                  env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_one_core_child_mode():
    r = run_bench("--config", "cfg3", "--cpu-one-core", "1", env={"BF_ORACLE_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["events"] > 50000 and d["seconds"] > 0 and d["cores"] == 1


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


# ---- control flow of the product arm at N = 2, without GPUs ----------------------------------------------------
# The CUDA pieces (streams, events, the context behind the C ABI, device tensors) are replaced by CPU stand-ins and
# NCCL by gloo; what runs for real is bench.main(): argument handling, batch assembly, warm-up, the timed loops, the
# per-step record snapshots and the ONE all_gather per timed region (shard.RecordGather), max-over-ranks timing, and
# the one JSON line on rank 0 only.

def _stub_worker(rank, world, port, q, break_gather):
    import contextlib
    import time
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import torch.distributed as dist
    import better_flow_b200 as bf
    from better_flow_b200 import shard
    import bench

    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))

    class FakeStream:
        cuda_stream = 0

    class FakeEvent:
        def __init__(self, enable_timing=False):
            self.t = 0.0

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.Stream = FakeStream
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.Event = FakeEvent
    torch.cuda.synchronize = lambda *a, **k: None
    real_empty, real_tensor, real_as_tensor = torch.empty, torch.tensor, torch.as_tensor

    def on_cpu(fn):
        def wrapped(*a, **k):
            if k.get("device") is not None:
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapped
    torch.empty, torch.tensor = on_cpu(real_empty), on_cpu(real_tensor)

    calls = {"launch": 0, "streamed": 0, "gathers": 0}
    dealt = []

    class FakeContext:
        launches = 0

        def __init__(self, rows, cols, scale, max_events, max_slices, device=0):
            self.buf = np.zeros(max_events, dtype=bf.EVENT_DTYPE)
            self.sl = []
            self.rec = real_empty(0, dtype=torch.uint8)

        def set_option(self, k, v): pass
        def get_option(self, k): return {"group_size": 2, "n_groups": 148}.get(k, 0)
        def set_stream(self, s): pass
        def staging(self): return self.buf
        def reset(self): self.sl = []
        def add_staged(self, off, n, scale, max_iter):
            self.sl.append(n)
            dealt.append(self.buf[off:off + n].copy())
        def upload(self): pass
        def sync(self): pass
        def close(self): pass

        def _run(self):
            FakeContext.launches += 1
            self.launches = FakeContext.launches
            # this rank's records of this launch: a recognisable byte pattern
            self.rec = torch.full((len(self.sl) * bf.RESULT_BYTES,), (7 * rank + self.launches) % 251, dtype=torch.uint8)

        def launch(self, want): calls["launch"] += 1; self._run()
        def run_streamed(self, want): calls["streamed"] += 1; self._run()
        def results_device(self): return 0, len(self.sl) * bf.RESULT_BYTES

        def results(self):
            return [{"rc": 0, "iters": 10 + rank, "n_events": n, "img_rows": 543, "img_cols": 723, "model": np.zeros(11)} for n in self.sl]

    bf.Context = FakeContext
    fake_ctx_holder = {}
    orig_init = FakeContext.__init__

    def init(self, *a, **k):
        orig_init(self, *a, **k)
        fake_ctx_holder["ctx"] = self
    FakeContext.__init__ = init
    torch.as_tensor = lambda obj, **k: fake_ctx_holder["ctx"].rec if hasattr(obj, "__cuda_array_interface__") else real_as_tensor(obj)
    bench.DEV = "cpu"

    real_init_pg = dist.init_process_group
    dist.init_process_group = lambda backend, **k: real_init_pg("gloo", rank=rank, world_size=world)
    real_flush = shard.RecordGather.flush

    def flush(self):
        if break_gather:
            raise RuntimeError("simulated failure of the deferred gather")
        out = real_flush(self)
        if out is not None:
            calls["gathers"] += 1
            calls["last_shape"] = tuple(out.shape)
            calls["last_ok"] = all(bool((out[r] == out[r][0, 0]).all()) or True for r in range(world)) and \
                [int(out[r][-1, 0]) for r in range(world)] == [(7 * r + FakeContext.launches) % 251 for r in range(world)]
        return out
    shard.RecordGather.flush = flush

    lines = []
    bench.emit = lambda line: lines.append(line)
    bench.quiet_stdout = lambda: None
    sys.argv = ["bench.py", "--gpus", str(world), "--steps", "4", "--warmup", "3", "--slices", "3", "--cpu-sample", "0"]
    bench.main()
    # the pooled deal: this rank must hold pool slices shard.partition(2 * 3, 2, rank, 4) of the rank-major pool
    pool = []
    for r in range(world):
        pool += [bf.pack_events(s.fr_x, s.fr_y, s.t_ns) for s in bench.make_batch(100 + r, 3)]
    want = [pool[g] for g in shard.partition(world * 3, world, rank, 4)]
    calls["deal_ok"] = len(want) == len(dealt) and all(np.array_equal(a, b) for a, b in zip(want, dealt))
    calls["n_dealt"] = len(dealt)
    q.put((rank, lines, calls))


@pytest.mark.parametrize("break_gather", [False, True])
def test_product_arm_control_flow_two_ranks_with_stubs(break_gather):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_stub_worker, args=(r, 2, port, q, break_gather)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, (lines, calls)) for r, lines, calls in (q.get(timeout=300) for _ in procs))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(res[0][0]) == 1 and res[1][0] == []                      # rank 0 prints the one line, rank 1 nothing
    d = res[0][0][0]
    assert d["n_gpus"] == 2 and d["steps"] == 4 and d["warmup"] == 3 and d["gpu_launches"] == 4
    assert d["config"]["events_per_step_per_gpu"] > 80000 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["scaling"] == "weak" and "dealt block-cyclically" in d["config"]["partition"]
    assert d["gather_ok"] is True and [p["rank"] for p in d["per_rank"]] == [0, 1]
    assert [p["slices"] for p in d["per_rank"]] == [4, 2] and d["per_rank"][1]["sum_iters"] == 2 * 11
    assert res[0][1]["deal_ok"] and res[1][1]["deal_ok"]
    for r in (0, 1):
        calls = res[r][1]
        assert calls["streamed"] == 5                                   # one warm e2e step + 4 timed
        if break_gather:
            assert calls["launch"] == 3 + 3 + 4 and calls["gathers"] == 0      # warm-up repeated on the fallback path
            assert "per step" in d["config"]["collective"]
        else:
            assert calls["launch"] == 3 + 4
            assert calls["gathers"] == 4                                # warm-up, timed resident, warm e2e, timed e2e
            assert calls["last_shape"] == (2, 4, 4 * 160) and calls["last_ok"]
            assert "4 timed steps" in d["config"]["collective"]
