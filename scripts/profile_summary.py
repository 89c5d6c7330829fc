"""Turn an `ncu --set full` report into the markdown summary committed under profiles/.
usage: python scripts/profile_summary.py REPORT.ncu-rep > profiles/NAME.md"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->L1 bytes"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
]
print(f"# ncu summary of `{rep.split('/')[-1]}`\n")
print("Captured with `ncu --set full --clock-control none --import-source on` on one B200 (sm_100a); durations are")
print("under the profiler (serialised, cold cache) -- bench.py's CUDA-event numbers are the ones to quote.\n")
for launch_no, r in enumerate(rows[2:]):
    name = r[idx["Kernel Name"]]
    print(f"## {name}\n")
    print("| metric | value |\n|---|---|")
    for m, label in METRICS:
        if m in idx:
            print(f"| {label} (`{m}`) | {r[idx[m]]} {units[idx[m]]} |")
    print()
    short = name.split("(")[0].split("::")[-1].split("<")[0]
    for view, title in (("cuda", "Hottest source lines (warp-stall samples)"),):
        # select by position in the report: template instances of one kernel share a name prefix, and a name filter
        # would return the first instance's source page for all of them
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                              "--launch-skip", str(launch_no), "--launch-count", "1"], capture_output=True, text=True).stdout
        rr = list(csv.reader(io.StringIO(out)))
        h2, cur, lines = None, "", []
        ops = collections.Counter()
        for q in rr:
            if len(q) == 2 and q[0] == "File Path":
                cur = q[1].split("/")[-1]
            elif len(q) > 10 and q[0] == "Line No":
                h2 = q
            elif h2 and len(q) == len(h2):
                try:
                    smp, ins = int(q[h2.index("# Samples")]), int(q[h2.index("Instructions Executed")])
                except ValueError:
                    continue
                if q[0] not in ("", "Line No"):
                    d = dict(zip(h2, q))
                    st = sorted(((k, int(d[k])) for k in h2 if k.startswith("stall_") and "(Not" not in k and d[k].isdigit()), key=lambda kv: -kv[1])[:2]
                    lines.append((smp, ins, cur, q[0], q[1].strip()[:80], st))
                elif q[2] not in ("", "-", "..."):
                    txt = q[3].strip()
                    op = (txt.split()[1] if txt.startswith("@") else txt.split()[0]).split(".")[0] if txt else "?"
                    ops[op] += ins
        ts, ti = sum(l[0] for l in lines) or 1, sum(l[1] for l in lines) or 1
        print(f"{title}:\n\n| samples % | instr % | line | source | top stalls |\n|---|---|---|---|---|")
        for l in sorted(lines, key=lambda l: -l[0])[:8]:
            print(f"| {100*l[0]/ts:.1f} | {100*l[1]/ti:.1f} | {l[2]}:{l[3]} | `{l[4]}` | {', '.join(f'{k} {v}' for k, v in l[5])} |")
        to = sum(ops.values()) or 1
        print("\nSASS opcode mix: " + ", ".join(f"{k} {100*v/to:.1f}%" for k, v in ops.most_common(10)) + "\n")
