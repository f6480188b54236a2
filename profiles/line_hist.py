"""Per-source-line dynamic instruction / stall-sample profile of one kernel.

Joins the SASS page of an ncu report (instructions executed + warp-stall samples per SASS instruction) with the line table
of the cubin (nvdisasm -gi), attributing every instruction to the OUTERMOST line inside the kernel's own file (inlined
device functions are charged to their call site).

    cuobjdump -xelf all lib.so ; nvdisasm -gi -c dcr_front.sm_100a.cubin > front.dis
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > k.csv
    python profiles/line_hist.py k.csv front.dis MANGLED_NAME SOURCE.cu [npx] [top]
"""
import csv
import re
import sys
from collections import Counter

kcsv, dis, mangled, srcfile = sys.argv[1:5]
npx = float(sys.argv[5]) if len(sys.argv) > 5 else None
top = int(sys.argv[6]) if len(sys.argv) > 6 else 60

# line table: instruction index -> outermost line in srcfile
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + mangled + ":"))
cur = None
table = []
pat_file = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = pat_file.search(l)
    if m:
        if m.group(3):
            cur = (m.group(3), int(m.group(4)))   # nvdisasm prints the outermost call site last on the chain line
        else:
            cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/", l):
        table.append(cur)

rows = list(csv.reader(open(kcsv)))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name")
hdr = rows[k + 1]
body = []
for r in rows[k + 2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        body.append(r)
ti, si = hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
assert len(body) == len(table), (len(body), len(table))
ins, smp = Counter(), Counter()
for r, loc in zip(body, table):
    key = loc[1] if loc and loc[0].endswith(srcfile.split("/")[-1]) else -1
    ins[key] += int(r[ti]); smp[key] += int(r[si])
tot, stot = sum(ins.values()), sum(smp.values())
src = open(srcfile).read().split("\n")
print("thread-instructions: %d%s; stall samples: %d" % (tot, (" = %.1f / px" % (tot / npx)) if npx else "", stot))
print("%6s %8s %7s %7s  source" % ("line", "instr/px" if npx else "instr%", "instr%", "stall%"))
for key, c in sorted(ins.items(), key=lambda kv: -kv[1])[:top]:
    text = src[key - 1].strip()[:100] if key > 0 else "(other files)"
    print("%6d %8.2f %6.1f%% %6.1f%%  %s" % (key, (c / npx) if npx else 100.0 * c / tot, 100.0 * c / tot, 100.0 * smp[key] / max(stot, 1), text))
