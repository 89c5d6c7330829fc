"""Top SASS instructions by executed count for a kernel in an ncu report (source page, sass view)."""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                      "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
ops = collections.Counter(); smp = collections.Counter()
tot = 0
for r in rows:
    if len(r) > 10 and r[0] in ("Address", "Line No"):
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        try:
            inst = int(r[hdr.index("Instructions Executed")]); s = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        src = r[hdr.index("Source")].strip()
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        op = op.split(".")[0]
        ops[op] += inst; smp[op] += s; tot += inst
print("total", tot)
for op, n in ops.most_common(25):
    print(f"{op:12s} {n:14d} {100*n/tot:5.1f}%  samples {smp[op]}")
