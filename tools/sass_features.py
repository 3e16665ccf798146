"""What the minimise kernel's instances are made of, straight from the SASS of a built library (no GPU needed):
instruction counts, spills, barriers, atomics / reductions, L2-coherent loads and the sm_100-specific instructions
(thread-block-cluster barrier UCGABAR_*, MAPA + cluster-window loads, and -- in a -DBF_TMA_PATCH=1 build -- UTMALDG /
SYNCS), with the lines around the first occurrence of each as an excerpt.
usage: sass_features.py lib.so [label]"""
import collections, re, subprocess, sys
lib = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else lib
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
KEY = re.compile(r"^(UCGABAR_ARV|UCGABAR_WAIT|UTMALDG|UTMASTG|UTMAPF|UBLKCP|SYNCS|MAPA|CCTL|MEMBAR|REDG|REDS|ATOMG|ATOMS|ATOM|ERRBAR|FENCE|STL|LDL|BAR|NANOSLEEP|LDG|LD|STG|LDS|STS|DADD|DMUL|DFMA|F2F|I2F|F2I|MUFU)")
EXCERPT = ("UCGABAR_ARV", "UCGABAR_WAIT", "MAPA", "UTMALDG", "SYNCS", "REDG", "CCTL")
print("== %s" % label)
for p in parts[1:]:
    name = p.split("\n", 1)[0].strip()
    if "bf_minimize_kernel" not in name:
        continue
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    lines = []
    for line in p.split("\n"):
        m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            lines.append((m.group(1), m.group(2).strip()))
    c = collections.Counter()
    first = {}
    for k, (addr, ins) in enumerate(lines):
        op = re.sub(r"^@!?U?P\w+\s+", "", ins).split()[0]
        m = KEY.match(op)
        if not m:
            continue
        fam = m.group(1)
        name_k = op if fam in ("UCGABAR_ARV", "UCGABAR_WAIT", "UTMALDG", "SYNCS", "MEMBAR", "CCTL", "REDG", "ATOMG", "MAPA", "ERRBAR") else fam
        if fam in ("LDG", "LD") and "STRONG" in op:
            name_k = fam + " ... STRONG.GPU"
        c[name_k] += 1
        for e in EXCERPT:
            if op.startswith(e) and e not in first:
                first[e] = k
    print("\n-- %s: %d SASS instructions" % (re.sub(r"\(KParams.*", "", dem), len(lines)))
    print("   " + ", ".join("%s %d" % kv for kv in sorted(c.items())))
    for e in EXCERPT:
        if e in first and e in ("UCGABAR_ARV", "UCGABAR_WAIT", "MAPA", "UTMALDG", "SYNCS"):
            k = first[e]
            print("   excerpt around the first %s:" % e)
            for addr, ins in lines[max(0, k - 3):k + 4]:
                print("      /*%s*/  %s" % (addr, ins))
    # distributed shared memory: `mapa` + ld.shared::cluster are lowered to a PRMT of the peer's rank into the address,
    # the shared-window base SR_SWINHI as the high word, and a generic 64-bit load
    sw = [k for k, (a, i) in enumerate(lines) if "SR_SWINHI" in i and "PRMT" in " ".join(x[1] for x in lines[k:k + 6])]
    if sw:
        k = sw[0]
        ld = next((j for j in range(k, min(len(lines), k + 80)) if re.match(r"^(@!?U?P\w+\s+)?LD\.E\.64", lines[j][1])), None)
        print("   excerpt: a peer CTA's partial sums read through the cluster's shared-memory window (mapa + ld.shared::cluster):")
        for addr, ins in lines[k:k + 6] + ([("...", "...")] + lines[ld - 1:ld + 2] if ld else []):
            print("      /*%s*/  %s" % (addr, ins))
