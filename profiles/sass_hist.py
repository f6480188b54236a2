"""Dynamic per-opcode histogram + top stall sites of one kernel from an ncu report.

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > k.csv ; python profiles/sass_hist.py k.csv [npx]
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
npx = float(sys.argv[2]) if len(sys.argv) > 2 else None
# the file may hold several kernels: "Kernel Name" line, header line, then SASS lines
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == "Kernel Name":
        name = rows[k][1]
        hdr = rows[k + 1]
        j = k + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(rows[j])
            j += 1
        k = j
        ci = hdr.index("Instructions Executed")
        ti = hdr.index("Thread Instructions Executed")
        si = hdr.index("# Samples")
        src = hdr.index("Source")
        ops, tops, samp = Counter(), Counter(), Counter()
        tot = ttot = stot = 0
        for r in body:
            op = r[src].split()
            if not op:
                continue
            o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
            o = o.split(".")[0] + ("." + o.split(".")[1] if o.startswith(("F2F", "I2F", "F2I", "MUFU", "LDS", "STS", "LDG", "STG", "ATOM", "RED")) and "." in o else "")
            c, t, s = int(r[ci]), int(r[ti]), int(r[si])
            ops[o] += c; tops[o] += t; samp[o] += s
            tot += c; ttot += t; stot += s
        print("==", name, "static SASS:", len(body), "warp-instr:", tot, "thread-instr:", ttot,
              ("= %.1f thread-instr/px" % (ttot / npx)) if npx else "")
        for o, c in ops.most_common(28):
            print("  %-14s %11d %5.1f%%   stall samples %5.1f%%" % (o, c, 100.0 * c / tot, 100.0 * samp[o] / max(stot, 1)))
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = Counter()
        for r in body:
            for h in stalls:
                agg[h] += int(r[hdr.index(h)])
        tots = sum(agg.values())
        print("  stall reasons (all samples):", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / tots) for h, v in agg.most_common(8)))
        top = sorted(body, key=lambda r: -int(r[si]))[:14]
        for r in top:
            print("   %5.2f%%  %s" % (100.0 * int(r[si]) / max(stot, 1), r[src].strip()[:90]))
    else:
        k += 1
