"""Local-memory (spill) instructions of a kernel by CUDA source line. usage: sass_spills.py lib.so kernel_substring"""
import re, collections, subprocess, sys, tempfile, os
lib, kname = os.path.abspath(sys.argv[1]), sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cnt = collections.Counter(); tot = 0; cur = None; ink = False
for cubin in sorted(os.listdir(tmp)):
    for l in subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines():
        m = re.match(r"^\.text\.(\S+):", l)
        if m: ink = kname in m.group(1); continue
        if not ink: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip()[:60]); continue
        if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", l):
            tot += 1
            m = re.search(r"\b(STL|LDL)(\.\S+)?\b", l)
            if m: cnt[(cur, m.group(1))] += 1
print(tot, "instructions; local-memory ops:", sum(cnt.values()))
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 50]: print(v, k)
