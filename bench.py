#!/usr/bin/env python
"""bench.py -- Mevents/s motion-compensated on B200 (BASELINE.json metric), next to the reference's
own CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4|cfg5]
                  [--scaling weak|strong]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: every slice of the batch
is minimised (OptimizerRolling::run, GD to convergence) by ONE persistent kernel launch.
Workload = BASELINE.json configs[1]: DAVIS-240C 240x180, 30 ms slices, GD to convergence,
stm-disabled (independent slices), synthetic 3 Mev/s contour stream (better_flow_b200/synth.py);
--config selects the other BASELINE.json configurations that fit one GPU (CONFIGS below).

  value     whole-job Mevents/s with the batch resident in HBM (CUDA events on the launch stream)
  e2e       same through the C ABI with the events in pinned HOST memory: H2D of the events and the
            slice table + launch + D2H of the per-slice results inside the timed region
  roofline  algorithmic bytes of SURVEY.md 8(d): A = sum_slices iters * (40 N + 16 P), divided by
            the launch duration, against MEASURED_PEAKS.json:hbm_gbs
  cpu_baseline  the reference's unmodified C++ (oracle/_ref, "reference") or its C restatement
            ("port") timed on this box's host cores on a bounded sample of the same slices: all
            threads (value), one pinned core (value_1core)
  parity    the GPU results of the timed batch against the reference's own arithmetic (oracle/_ref, else the
            C restatement) on a sample of the same slices: max relative deviation of (total_dx, total_dy),
            iteration counts equal or not -- the number carries its correctness bit
At N > 1 (weak scaling) the ranks' streams form ONE pool of N x S slices that is dealt block-cyclically
(shard.partition, blocks of 4: SURVEY 8e; shard.deal_pool) so that every rank minimises S slices drawn evenly
from all streams, with no data-path collective; the per-slice flow records of all timed steps are gathered with
ONE NCCL all_gather inside the timed region (each step leaves a device-side snapshot of its records;
better_flow_b200/shard.py: RecordGather) and every rank checks its slot of the gathered tensor against its own
records (gather_ok).  --scaling strong shards ONE fixed stream of the configuration's slices over the ranks
(BASELINE.json configs[3], [4]).
stdout carries exactly one JSON line; everything else (library banners included) goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# BASELINE.json configurations that fit one GPU.  The default, cfg2, is the one the metric is
# quoted on (configs[1]); the others are selectable with --config for the multi-GPU runs BASELINE.json names
# (cfg4: slice-sharded over 2/4/8 GPUs, cfg5: 8 GPUs + one NCCL gather) -- same step, same JSON line.
#   cols, rows, events/s, slice length, slices per step per GPU, max_iter, label
CONFIGS = {
    # 4 slices per CTA group (148 groups of 2); 592 x ~90 k events x 8 B = 426 MB of events > 126 MB L2
    "cfg2": (240, 180, 3e6, 0.030, 592, -1, "DAVIS-240C 240x180 synthetic 3 Mev/s contour stream, 30 ms slices (~90k events)"),
    "cfg3": (346, 260, 2e6, 0.050, 64, -1, "DAVIS-346 346x260 synthetic 2 Mev/s contour stream, 50 ms slices (~100k events), 64 slices batched per launch"),
    "cfg4": (640, 480, 10e6, 0.020, 128, -1, "640x480 synthetic 10 Mev/s contour stream, 20 ms slices (200k events)"),
    "cfg5": (1280, 720, 100e6, 0.010, 32, -1, "1280x720 synthetic 100 Mev/s contour stream, 10 ms slices (1M events)"),
}
SCALE = 3
DEV = "cuda"   # (tests/test_bench_cpu.py runs main() with CPU stand-ins and sets this to "cpu")
METRIC = "Mevents/sec motion-compensated"
UNIT = "Mevents/s"


def select_config(name, slices=None):
    global SENSOR_COLS, SENSOR_ROWS, RATE_EPS, SLICE_S, SLICES_PER_STEP, MAX_ITER, WORKLOAD, CONFIG_NAME
    SENSOR_COLS, SENSOR_ROWS, RATE_EPS, SLICE_S, SLICES_PER_STEP, MAX_ITER, label = CONFIGS[name]
    if slices:
        SLICES_PER_STEP = slices
    CONFIG_NAME = name
    WORKLOAD = "%s, GD to convergence, scale 3, stm-disabled, %d slices per step" % (label, SLICES_PER_STEP)


select_config("cfg2")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def make_batch(seed, n_slices):
    from better_flow_b200 import synth
    st = synth.make_stream(SENSOR_COLS, SENSOR_ROWS, RATE_EPS, SLICE_S * n_slices, seed=seed)
    return synth.cut_slices(st, SLICE_S)[:n_slices]


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [v for v in sm if v > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(slices, max_seconds=None):
    """Time the reference CPU path (run() only, steady_clock inside the driver) on `slices`; the per-slice models
    are kept so that the caller can compare them with the GPU results of the same slices (parity block)."""
    os.environ.setdefault("BF_ORACLE_THREADS", str(os.cpu_count() or 1))   # torchrun pins OMP_NUM_THREADS=1
    from oracle import ref, port
    use_ref = ref.available(SENSOR_ROWS, SENSOR_COLS)
    ev = 0
    secs = 0.0
    iters = []
    models = []
    t_wall = time.perf_counter()
    done = 0
    for s in slices:
        if use_ref:
            r = ref.minimize(s.fr_x, s.fr_y, s.t_ns, scale=SCALE, max_iter=MAX_ITER, rows=SENSOR_ROWS, cols=SENSOR_COLS)
            secs += r["seconds"]
        else:
            t0 = time.perf_counter()
            r = port.minimize(s.fr_x, s.fr_y, s.t_ns, scale=SCALE, max_iter=MAX_ITER, rows=SENSOR_ROWS, cols=SENSOR_COLS)
            secs += time.perf_counter() - t0
        ev += len(s.fr_x)
        iters.append(r["iters"])
        models.append(np.array(r["model"], dtype=np.float64))
        done += 1
        if max_seconds is not None and time.perf_counter() - t_wall > max_seconds:
            break
    cores = ref.load(SENSOR_ROWS, SENSOR_COLS).threads if use_ref else 1
    return {"events": ev, "seconds": secs, "slices": done, "iters_mean": float(np.mean(iters)), "iters": iters, "models": models,
            "kind": "reference" if use_ref else "port", "cores": cores,
            # oracle/shim/tbb/parallel_for.h: the reference's two tbb::parallel_for row loops run on an OpenMP stand-in,
            # not on the vendored TBB (results are thread-count independent)
            "threads_impl": "openmp-shim" if use_ref else "serial-port"}


def parity_block(gpu_results, cpu_info):
    """GPU result records of the timed batch vs the reference's models of the same slices (first len(models) slices):
    the contract is 1e-4 relative on (total_dx, total_dy), model[7:9]."""
    n = len(cpu_info["models"])
    rels, it_eq = [], True
    for r, m, it in zip(gpu_results[:n], cpu_info["models"], cpu_info["iters"]):
        g = np.asarray(r["model"], dtype=np.float64)
        rels.append(float(np.max(np.abs(g[7:9] - m[7:9]) / np.maximum(np.abs(m[7:9]), 1e-300))))
        it_eq = it_eq and int(r["iters"]) == int(it)
    mx = max(rels) if rels else None
    return {"n": n, "max_rel_dxdy": mx, "iters_equal": bool(it_eq), "tolerance": 1e-4, "against": cpu_info["kind"],
            "ok": bool(rels and mx < 1e-4)}


class _Sl:
    """A slice given as packed 8-byte records (what a rank holds after the pooled deal)."""
    def __init__(self, ev):
        self.ev = ev
        self.fr_x = np.ascontiguousarray(ev["fr_x"])
        self.fr_y = np.ascontiguousarray(ev["fr_y"] & 0x7fff)
        self.t_ns = np.ascontiguousarray(ev["t_ns"])


def class_surface_e2e(device_index):
    """The reference-facing CLASS surface, end to end: better_flow_b200/bf_motion_compensator -- the
    drop-in tool, i.e. Event construction + DVS_flow::add_event per event, the ring buffer, the triggers, one
    OptimizerRolling minimisation per overlapping 50 k-event window warm-started from the previous one (the reference's
    DEFAULT mode) -- on a 4.5 M-event DAVIS-240C stream that is already in memory (the tool's --bufferize-file
    semantics), timed by the tool itself with steady_clock around the whole add_event loop + final slice + model
    read-back.  Returns new events per second, or None when the tool is not built."""
    import re
    import tempfile
    cli = os.path.join(ROOT, "better_flow_b200", "bf_motion_compensator")
    if not os.path.exists(cli):
        return None
    try:
        from better_flow_b200 import synth
        st = synth.make_stream(240, 180, 3e6, 1.5, seed=1)
        rec = np.zeros(len(st), dtype=np.dtype([("t", "<u8"), ("x", "<u2"), ("y", "<u2"), ("p", "<u4")]))
        rec["t"], rec["x"], rec["y"], rec["p"] = st.t_ns, st.x, st.y, st.p
        with tempfile.TemporaryDirectory() as d:
            binf = os.path.join(d, "stream.bin")
            rec.tofile(binf)
            out = {}
            for key, extra in (("default_mode", []), ("default_mode_host_ring", ["--no-device-ring"])):
                best = None
                for _ in range(2):
                    r = subprocess.run([cli, "--quiet", "--device=%d" % device_index] + extra + [binf], capture_output=True, text=True,
                                       env=dict(os.environ, BF_TIMING="1"), timeout=45)   # (a TimeoutExpired abandons the whole extra)
                    m = re.search(r"\[timing\] processing (\d+) events in ([0-9.e+-]+) s = ([0-9.e+-]+) Mev/s, slices (\d+)", r.stderr)
                    if r.returncode == 0 and m and (best is None or float(m.group(2)) < best[0]):
                        best = (float(m.group(2)), float(m.group(3)), int(m.group(4)))
                if best:
                    out[key] = {"value": best[1], "seconds": best[0], "slices": best[2]}
        if "default_mode" not in out:
            return None
        return {"value": out["default_mode"]["value"], "unit": UNIT, "events": len(st), "modes": out,
                "what": "bf_motion_compensator (DVS_flow class surface), reference default mode: overlapping 50 k-event / 200 ms windows every "
                        "20 k events, warm-start chain, GD to convergence; events in host memory, device context and slice ring created before the stream "
                        "(DVS_flow::prepare), add_event loop + all slices + model read-back timed"}
    except Exception as exc:   # the bench line must not depend on this extra
        print("class-surface measurement skipped: %s" % exc, file=sys.stderr)
        return None


def bind_to_gpu_numa_node(torch, index):
    """N > 1: run this rank (and so allocate its pinned staging buffer, first touch) on the CPUs of the NUMA node
    its GPU hangs off; 8 ranks streaming 426 MB per step each otherwise cross the socket link.  Best effort."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            return
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) >= 2:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def one_core_baseline(n_slices):
    """The same CPU path pinned to ONE core (SURVEY 8d asks for both figures): a child process, because the
    thread count of the compiled reference is fixed at its first parallel_for."""
    cmd = [sys.executable, os.path.abspath(__file__), "--config", CONFIG_NAME, "--cpu-one-core", str(n_slices)]
    try:
        if subprocess.run(["taskset", "-c", "0", "true"], capture_output=True).returncode == 0:
            cmd = ["taskset", "-c", "0"] + cmd
        env = dict(os.environ, BF_ORACLE_THREADS="1", OMP_NUM_THREADS="1")
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
        r = json.loads(out)
        return r["events"] / r["seconds"] / 1e6
    except Exception:
        return None


def slice_parallel_baseline(n_slices):
    """What the host could do with one single-threaded reference process per core, each on its own slices (the
    slices are independent under --stm-disable; the reference's own tool is one sequential process, so this is a
    harness around it, reported next to the figure of the reference as it runs).  Children as in one_core_baseline;
    throughput = all events / the slowest child's run() time."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        cmd = [sys.executable, os.path.abspath(__file__), "--config", CONFIG_NAME, "--cpu-one-core", str(n_slices)]
        pin = subprocess.run(["taskset", "-c", str(cpus[0]), "true"], capture_output=True).returncode == 0
        env = dict(os.environ, BF_ORACLE_THREADS="1", OMP_NUM_THREADS="1")
        procs = [subprocess.Popen((["taskset", "-c", str(c)] if pin else []) + cmd, env=env, stdout=subprocess.PIPE,
                                  stderr=subprocess.DEVNULL, text=True) for c in cpus]
        outs = [json.loads(p.communicate(timeout=300)[0].strip().splitlines()[-1]) for p in procs]
        return {"value": sum(o["events"] for o in outs) / max(o["seconds"] for o in outs) / 1e6, "processes": len(outs)}
    except Exception:
        return None


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation on the host cores, bounded sample per step."""
    if rank != 0:
        return
    per_step = {"cfg2": 6, "cfg3": 4, "cfg4": 2, "cfg5": 1}[CONFIG_NAME]   # ~0.2-3 s of CPU work per slice
    # the FIRST slices of the product arm's own batch (same generator call, seed 100), so both arms see the same input
    need = per_step * (args.steps + args.warmup)
    slices = make_batch(100, max(SLICES_PER_STEP, need))[:need]
    k = 0
    for _ in range(args.warmup):
        cpu_reference_run(slices[k:k + per_step]); k += per_step
    ev, secs, its = 0, 0.0, []
    info = None
    for _ in range(args.steps):
        info = cpu_reference_run(slices[k:k + per_step]); k += per_step
        ev += info["events"]; secs += info["seconds"]; its.append(info["iters_mean"])
    val = ev / secs / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "name": CONFIG_NAME,
                   "sample": "%d slices per step: the first slices of the product arm's batch (seed 100)" % per_step,
                   "iters_mean": float(np.mean(its))},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                         "threads_impl": info["threads_impl"],
                         "sample": "%d steps x %d slices, OptimizerRolling::run() wall time only" % (args.steps, per_step),
                         "slice_parallel": slice_parallel_baseline(max(1, per_step // 2))},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_JSON_OUT = None


def quiet_stdout():
    """Everything any library prints to stdout from here on (NCCL's version banner, ...) goes to stderr; the one
    JSON line is written to the original stdout by emit()."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE.json configuration (default: the one the metric is quoted on)")
    ap.add_argument("--slices", type=int, default=0, help="slices per step per GPU (0 = the configuration's)")
    ap.add_argument("--group-size", type=int, default=0)
    ap.add_argument("--upload-chunks", type=int, default=0, help="chunks of the streamed event upload (0 = library default)")
    ap.add_argument("--cpu-sample", type=int, default=24, help="slices timed on the CPU baseline (0 = skip)")
    ap.add_argument("--cpu-one-core", type=int, default=0, help=argparse.SUPPRESS)   # child mode of one_core_baseline()
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: S slices per GPU (pooled block-cyclic deal at N > 1); strong: ONE stream of S slices sharded over the GPUs")
    ap.add_argument("--upload-format", default="plain", choices=["plain", "delta"],
                    help="end-to-end path: 8-byte records (default) or 6-byte delta records (25 %% fewer H2D bytes; at 8 GPUs, where four of them "
                         "share one host link, 28.6 instead of 22.7 Gev/s end to end: profiles/r2m_*).  Opt-in: with the 1280x720 configuration on 4 "
                         "GPUs two of three runs stalled in the delta instance of the kernel (not reproduced on one GPU, unexplained: DESIGN.md 7)")
    ap.add_argument("--opt", default="", help="library options key=value[,key=value] (development: A/B of kernel variants)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    select_config(args.config, args.slices)
    args.slices = SLICES_PER_STEP

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.cpu_one_core > 0:
        info = cpu_reference_run(make_batch(100, args.cpu_one_core))
        emit({"events": info["events"], "seconds": info["seconds"], "cores": info["cores"], "iters_mean": info["iters_mean"]})
        return
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import better_flow_b200 as bf

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_to_gpu_numa_node(torch, local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # (a short collective timeout: should a rank ever stall, the run must fail within minutes, not after NCCL's default 10)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))

    import better_flow_b200.shard as shard

    # ---- this rank's slices -------------------------------------------------------------------------------
    # weak scaling : every rank generates its own stream of S slices (seed 100 + rank); at N > 1 the N streams form
    #                ONE pool of N x S slices dealt block-cyclically (shard.deal_pool = shard.partition, blocks of 4)
    # strong scaling: ONE stream of S slices (seed 100, generated identically on every rank), rank r takes
    #                shard.partition(S, N, r, 4)
    strong = args.scaling == "strong"
    if strong:
        pool = make_batch(100, args.slices)
        ids = shard.partition(len(pool), world, rank, 4)
        mine = [bf.pack_events(pool[i].fr_x, pool[i].fr_y, pool[i].t_ns) for i in ids]
        del pool
    else:
        local = make_batch(100 + rank, args.slices)
        mine = [bf.pack_events(s.fr_x, s.fr_y, s.t_ns) for s in local]
        ids = list(range(len(mine)))
        del local
        if dist is not None:
            ids, mine = shard.deal_pool(mine, dist, torch, DEV, block=4)
    n_events = int(sum(len(e) for e in mine))
    ctx = bf.Context(SENSOR_ROWS, SENSOR_COLS, SCALE, max_events=n_events + 64, max_slices=len(mine) + 1,
                     device=local_rank)
    if args.group_size:
        ctx.set_option("group_size", args.group_size)
    if args.upload_chunks:
        ctx.set_option("upload_chunks", args.upload_chunks)
    for kv in (args.opt or "").split(","):
        if kv:
            ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    # assemble the batch in the library's pinned staging: the compact upload format (6-byte delta records, expanded to
    # the kernel's 8-byte records on the device; include/bf_cuda.h) when every slice can be represented in it, else
    # plain 8-byte records
    ctx.reset()
    upload_format = "6-byte delta records"
    try:
        if args.upload_format == "plain":
            raise bf.BfError("plain upload requested")
        for e in mine:
            ctx.add_delta(e, SCALE, MAX_ITER)
    except (bf.BfError, AttributeError):
        upload_format = "8-byte records"
        stage = ctx.staging()
        off = 0
        ctx.reset()
        for e in mine:
            n = len(e)
            stage[off:off + n] = e
            ctx.add_staged(off, n, SCALE, MAX_ITER)
            off += n
    try:
        h2d = ctx.upload_bytes             # what one streamed run copies host -> device: event records (+ block table) + slice table
    except AttributeError:
        h2d = n_events * 8 + len(mine) * 120
    d2h = len(mine) * bf.RESULT_BYTES

    # N > 1: every step leaves a device-side snapshot of its per-slice flow records (a stream-ordered ~100 KB D2D
    # copy), and ONE NCCL all_gather inside the timed region exchanges the records of all its steps
    # (shard.RecordGather; SURVEY 8e: "one NCCL gather of the resulting flow vectors").
    rg = None
    rec_cap = 0
    if dist is not None:
        t = torch.tensor([len(mine)], device=DEV, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rec_cap = int(t.item()) * bf.RESULT_BYTES                    # (strong scaling: shares differ by up to one block)
        rg = shard.RecordGather(dist, torch, rec_cap, max(args.steps, args.warmup, 1) + 1, DEV)

    class _Dev:  # __cuda_array_interface__ view of the device result records
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    fallback = {"buf": None}   # set if the deferred gather cannot be used: one in-line all_gather per step instead
    pad = {"buf": None}

    def my_records():
        ptr, nbytes = ctx.results_device()
        t = torch.as_tensor(_Dev(ptr, nbytes), device=DEV)
        if nbytes == rec_cap:
            return t
        if pad["buf"] is None:
            pad["buf"] = torch.zeros(rec_cap, dtype=torch.uint8, device=DEV)
        pad["buf"][:nbytes].copy_(t, non_blocking=True)
        return pad["buf"]

    def gather():
        if dist is None:
            return
        if fallback["buf"] is not None:
            dist.all_gather_into_tensor(fallback["buf"], my_records())
            return
        rg.snapshot(my_records())                     # on the launch stream, before the next launch

    def flush_gather():
        if rg is not None and fallback["buf"] is None:
            rg.flush()

    def step_resident():
        ctx.launch(False)
        gather()

    def step_e2e():
        # H2D of the events (streamed in slice-ordered chunks into the device buffer the PREVIOUS launch is not
        # reading, so it also overlaps that launch) + the one persistent launch + D2H of the result records
        ctx.run_streamed(False)
        gather()

    def timed(fn, k):
        """-> (max over ranks, this rank's own) milliseconds for k steps."""
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, em = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(k):
                fn()
            em.record(stream)                         # this rank's own work ends here; the collective below waits for every rank
            flush_gather()                            # the one all_gather of these k steps' records: inside the timed region
            e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        own = e0.elapsed_time(em)
        if dist is not None:
            t = torch.tensor([ms], device=DEV, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, own

    with torch.cuda.stream(stream):
        ctx.upload()
        try:
            for _ in range(args.warmup):
                step_resident()
            flush_gather()                            # (also creates the NCCL communicator outside the timed region)
        except Exception as exc:                      # same code and shapes on every rank: they all land here together
            if rg is None:
                raise
            print("deferred record gather unavailable (%s): one all_gather per step instead" % exc, file=sys.stderr)
            fallback["buf"] = torch.empty(world * rec_cap, dtype=torch.uint8, device=DEV)
            for _ in range(args.warmup):
                step_resident()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = ctx.launches
    ms_res, own_res = timed(step_resident, args.steps)
    launches = ctx.launches - launches0
    with torch.cuda.stream(stream):
        step_e2e()
        flush_gather()
    ms_e2e, own_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None
    ctx.sync()

    res = ctx.results()
    iters = [r["iters"] for r in res]
    ok = all(r["rc"] == 0 for r in res)
    P = res[0]["img_rows"] * res[0]["img_cols"]
    alg_bytes = float(sum(r["iters"] * (40 * r["n_events"] + 16 * r["img_rows"] * r["img_cols"]) for r in res))
    alg_ev_bytes = float(sum(r["iters"] * 40 * r["n_events"] for r in res))
    sum_iters = int(sum(iters))
    event_iters = int(sum(r["iters"] * r["n_events"] for r in res))

    # gather_ok: this rank's slot of the gathered tensor (last timed step) is byte-identical to its own device records
    gather_ok = None
    per_rank = None
    tot_events = n_events
    if dist is not None:
        okf = 1.0
        if rg is not None and rg.last is not None and fallback["buf"] is None:
            ptr, nbytes = ctx.results_device()
            own_rec = torch.as_tensor(_Dev(ptr, nbytes), device=DEV)
            okf = 1.0 if bool(torch.equal(rg.last[rank, -1, :nbytes], own_rec)) else 0.0
        elif fallback["buf"] is not None:
            ptr, nbytes = ctx.results_device()
            own_rec = torch.as_tensor(_Dev(ptr, nbytes), device=DEV)
            okf = 1.0 if bool(torch.equal(fallback["buf"].view(world, -1)[rank, :nbytes], own_rec)) else 0.0
        stats = torch.tensor([own_res / args.steps, own_e2e / args.steps, float(sum_iters), float(event_iters), float(n_events),
                              float(len(mine)), okf], device=DEV, dtype=torch.float64)
        allst = torch.empty((world, stats.numel()), device=DEV, dtype=torch.float64)
        dist.all_gather_into_tensor(allst.view(-1), stats)
        allst = allst.cpu().numpy()
        tot_events = int(allst[:, 4].sum())
        gather_ok = bool(allst[:, 6].min() > 0.5)
        per_rank = [{"rank": r, "own_ms_per_step": float(allst[r, 0]), "own_e2e_ms_per_step": float(allst[r, 1]), "sum_iters": int(allst[r, 2]),
                     "event_iters": int(allst[r, 3]), "events": int(allst[r, 4]), "slices": int(allst[r, 5])} for r in range(world)]
    value = tot_events * args.steps / ms_res / 1e3
    e2e_val = tot_events * args.steps / ms_e2e / 1e3

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        ms_launch = own_res / args.steps              # rank 0's own launch duration: the roofline is this GPU's kernel
        achieved = alg_bytes / (ms_launch * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        cpu = None
        parity = None
        n_cpu = args.cpu_sample if world == 1 else min(args.cpu_sample, 8)   # (N > 1: a short parity sample only)
        if n_cpu > 0:
            info = cpu_reference_run([_Sl(e) for e in mine[:n_cpu]], max_seconds=40.0)
            parity = parity_block(res, info)
            if world == 1:
                cpu = {"value": info["events"] / info["seconds"] / 1e6, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                       "threads_impl": info["threads_impl"],
                       "value_1core": one_core_baseline({"cfg2": 12, "cfg3": 8, "cfg4": 3, "cfg5": 1}[CONFIG_NAME]),
                       "sample": "first %d slices of the same batch, OptimizerRolling::run() wall time only, iters mean %.1f"
                                 % (info["slices"], info["iters_mean"])}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "name": CONFIG_NAME,
                       "l2": ("inputs larger than L2 (%.0f MB of events per GPU per step)" % (n_events * 8 / 1e6)) if n_events * 8 > 130e6
                             else ("working set larger than L2 (%.0f MB of events + as much state + %.0f MB of slice images in flight per GPU)"
                                   % (n_events * 8 / 1e6, 16.0 * P * ctx.get_option("n_groups") / 1e6)),
                       "events_per_step_per_gpu": n_events, "pixels_per_image": P, "iters_mean": float(np.mean(iters)),
                       "iters_max": int(max(iters)), "all_converged": bool(ok), "group_size": ctx.get_option("group_size"),
                       "n_groups": ctx.get_option("n_groups"),
                       "partition": "one GPU" if world == 1 else
                                    ("ONE stream of %d slices sharded over the ranks with shard.partition (blocks of 4)" % args.slices if strong else
                                     "pool of %d x %d slices dealt block-cyclically over the ranks (shard.deal_pool, blocks of 4)" % (world, args.slices)),
                       "collective": "none" if world == 1 else
                                     ("one NCCL all_gather of per-slice flow records per step" if fallback["buf"] is not None else
                                      "one NCCL all_gather of the per-slice flow records of the %d timed steps (device snapshot per step), "
                                      "inside the timed region" % args.steps)},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "upload_format": upload_format,
                    "pipeline": "events re-uploaded every step from pinned host memory into the device buffer the previous launch is "
                                "not reading (two buffers), streamed in slice-ordered chunks the kernel consumes as they land"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "bf_minimize_kernel",
                         "algorithmic_bytes_per_launch": alg_bytes, "events_only_bytes_per_launch": alg_ev_bytes,
                         "launch_ms": ms_launch,
                         "note": "A = sum iters*(40N+16P), SURVEY 8(d); the image pass is sparse and the working set partly "
                                 "L2-resident, so measured DRAM traffic is below A (see profiles/)"},
            "clocks": clocks,
        }
        if world == 1 and CONFIG_NAME == "cfg2" and args.cpu_sample > 0:
            cs = class_surface_e2e(local_rank)
            if cs:
                line["e2e_class_surface"] = cs
        if parity:
            line["parity"] = parity
        if gather_ok is not None:
            line["gather_ok"] = gather_ok
            line["per_rank"] = per_rank
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
