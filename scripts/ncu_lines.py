"""Summarise `ncu --page source --csv` per CUDA source line: samples, instructions executed, top stall.
usage: python scripts/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [launch_skip] [top]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
cur_file = ""
lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = dict(zip(hdr, r))
        try:
            samples = int(r[hdr.index("# Samples")])
            inst = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        stalls = {k: int(d[k]) for k in hdr if k.startswith("stall_") and "(Not" not in k and d[k].isdigit()}
        ts = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        lines.append((samples, inst, cur_file, r[0], r[1].strip()[:90], ts))
tot_s = sum(l[0] for l in lines) or 1
tot_i = sum(l[1] for l in lines) or 1
print(f"total samples {tot_s} total warp-instr {tot_i}")
for l in sorted(lines, key=lambda l: -l[0])[:top]:
    print(f"{100*l[0]/tot_s:5.1f}% smp {100*l[1]/tot_i:5.1f}% inst  {l[2]}:{l[3]:>4}  {l[4]}   {l[5]}")
