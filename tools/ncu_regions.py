"""Warp-stall samples of the minimise kernel by REGION (event pass / image pass / barriers / control) and stall reason.
usage: ncu_regions.py sass.csv lib.so kernel_substring   (sass.csv = `ncu -i rep --page source --csv`)
Instructions are attributed to source lines with nvdisasm -g (like ncu_lines.py), lines to the device function that
contains them (function table read from bf_device.cuh / bf_cuda.cu), functions to regions."""
import csv, sys, subprocess, re, collections, os, tempfile
sass_csv, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REGION = {  # device function -> region
    "event": ["event_pass", "event_one", "project_event", "event_pixel", "pixel_offset", "mark_cells_bm", "stamp_bit", "flush_stamp_bitmap",
              "slope_time", "warp_from_m", "div_const", "u32_to_double", "make_pixel_map", "ld_nc_u32x4", "ld_state4", "st_state4", "red_add_u64",
              "local_event_pass"],
    "image": ["image_pass", "cell_process", "cell_compute", "local_cell_process_s5", "tma_issue_patch", "mbar_wait", "compact_cells", "cell_clear", "unpack_avg_fast", "local_cell_process", "acc_zero", "fill_rcp_table"],
    "reduce+GD": ["acc_block_reduce", "warp_sum", "group_sums", "group_sums_block_gather", "group_sums_block_finish", "opt_advance_warp",
                  "sincos_small", "local_opt_advance", "local_opt_init"],
    "barrier": ["group_barrier", "ld_acquire_u32", "ld_relaxed_u32", "red_release_add_u32", "fence_acq_rel_gpu"],
}
fn2region = {f: r for r, fs in REGION.items() for f in fs}

def function_table(path):
    """[(first_line, name)] of the __device__ / __global__ functions of a source file."""
    out = []
    src = open(path).read().splitlines()
    for i, l in enumerate(src, 1):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?__(?:device|global)__.*?\b([A-Za-z_0-9]+)\s*\(", l)
        if m and not l.rstrip().endswith(";"): out.append((i, m.group(1)))
    return out
tables = {f: function_table(os.path.join(ROOT, "better_flow_b200", "csrc", f)) for f in ("bf_device.cuh", "bf_cuda.cu")}

def region_of(f, ln):
    t = tables.get(f)
    if not t: return "intrinsics (shuffles, atomics)" if "intrinsics" in f or "atomic" in f else "other"
    name = None
    for first, n in t:
        if first <= ln: name = n
        else: break
    if name is None: return "other"
    return fn2region.get(name, "slice loop / control (" + ("bf_cuda.cu" if f == "bf_cuda.cu" else name) + ")")

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
lines = []; in_k = False; cur = ("?", 0)
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    for l in subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines():
        m = re.match(r"^\.text\.(\S+):", l)
        if m: in_k = kname in m.group(1); continue
        if not in_k: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
assert len(data) == len(lines), (len(data), len(lines))
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
for r, (f, ln) in zip(data, lines):
    reg = region_of(f, ln)
    for h in reasons:
        v = r[col[h]]
        if v: agg[reg][h] += float(v)
    agg[reg]["instr"] += float(r[col["Instructions Executed"]] or 0)
tot = sum(sum(v for k, v in c.items() if k != "instr") for c in agg.values())
ti = sum(c["instr"] for c in agg.values())
top = [h for h, _ in collections.Counter({h: sum(c[h] for c in agg.values()) for h in reasons}).most_common(9)]
print("stall samples by region (% of all samples); columns = the 9 most frequent reasons")
print("%-44s %7s %7s  " % ("region", "instr%", "smpl%") + " ".join("%9s" % h.replace("stall_", "")[:9] for h in top))
for reg, c in sorted(agg.items(), key=lambda kv: -sum(v for k, v in kv[1].items() if k != "instr")):
    s = sum(v for k, v in c.items() if k != "instr")
    print("%-44s %7.2f %7.2f  " % (reg[:44], 100 * c["instr"] / ti, 100 * s / tot) + " ".join("%9.2f" % (100 * c[h] / tot) for h in top))
print("%-44s %7.2f %7.2f  " % ("all", 100.0, 100.0) + " ".join("%9.2f" % (100 * sum(c[h] for c in agg.values()) / tot) for h in top))
