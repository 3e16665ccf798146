"""Attribute an `ncu --page source --csv` SASS dump to CUDA source lines using nvdisasm -g line info.
usage: ncu_lines.py sass.csv lib.so kernel_substring [topN]"""
import csv, sys, subprocess, re, collections, os, tempfile
sass_csv, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
dis = []
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# collect (file,line) per instruction of the kernel, in order
lines = []
in_k = False; cur = ("?", 0)
for l in dis:
    m = re.match(r"^\.text\.(\S+):", l)
    if m:
        in_k = kname in m.group(1); continue
    if not in_k: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
print("sass in csv", len(data), "sass in disasm", len(lines))
agg = collections.defaultdict(lambda: [0.0, 0.0])
n = min(len(data), len(lines))
for r, key in zip(data[:n], lines[:n]):
    try:
        agg[key][0] += float(r[col["Instructions Executed"]]); agg[key][1] += float(r[col["# Samples"]])
    except ValueError: pass
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
src = {}
def srcline(f, ln):
    if f not in src:
        p = [os.path.join(d, f) for d in ("better_flow_b200/csrc", ".") if os.path.exists(os.path.join(d, f))]
        src[f] = open(p[0]).read().splitlines() if p else []
    return src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
print("%-22s %8s %8s  %s" % ("file:line", "instr%", "samples%", "source"))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print("%-22s %8.2f %8.2f  %s" % ("%s:%d" % key, 100 * v[0] / ti, 100 * v[1] / ts, srcline(*key)))
