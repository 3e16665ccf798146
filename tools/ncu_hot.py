"""Summarise an `ncu --page source --csv` dump: hottest SASS regions by executed instructions / stall samples."""
import csv, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = rows[hdr_i + 1:]
def f(r, k):
    try: return float(r[col[k]])
    except Exception: return 0.0
tot_i = sum(f(r, "Instructions Executed") for r in data)
tot_s = sum(f(r, "# Samples") for r in data)
print("total warp-instr %.3e  samples %d  sass lines %d" % (tot_i, tot_s, len(data)))
# split into regions at branch targets / large count changes: here fixed windows of contiguous equal-count runs
regions = []
cur = None
for idx, r in enumerate(data):
    n = f(r, "Instructions Executed")
    if cur is None or abs(n - cur["n"]) > 0.05 * max(n, cur["n"], 1):
        cur = {"n": n, "start": idx, "end": idx, "instr": 0.0, "samples": 0.0}
        regions.append(cur)
    cur["end"] = idx; cur["instr"] += n; cur["samples"] += f(r, "# Samples")
regions.sort(key=lambda x: -x["samples"])
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for reg in regions[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    st = collections.Counter()
    for r in data[reg["start"]:reg["end"] + 1]:
        for s in stalls: st[s] += f(r, s)
    top = ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(reg["samples"], 1)) for k, v in st.most_common(4))
    print("\n== lines %d-%d  exec/line %.3e  instr %.1f%%  samples %.1f%%  [%s]" % (
        reg["start"], reg["end"], reg["n"], 100 * reg["instr"] / tot_i, 100 * reg["samples"] / tot_s, top))
    for r in data[reg["start"]:min(reg["end"] + 1, reg["start"] + 14)]:
        print("    %-70s %6.0f" % (r[col["Source"]][:70], f(r, "# Samples")))
