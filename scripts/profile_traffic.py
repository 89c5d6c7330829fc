"""DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the kernels in an `ncu --set full` report,
written to a json under profiles/ that bench.py reads for `roofline.traffic`.
usage: python scripts/profile_traffic.py REPORT.ncu-rep OUT.json WORKLOAD KEY [KEY ...]
KEYs name the launches of the report in order (e.g. qc_contract.fwd qc_dphi.bwd qc_contract.bwd for one fp32 step of
scripts/run_layer.py; qc_contract_mma.fwd qc_dphi_mma.bwd qc_dw_mma.bwd qc_contract_mma.bwd for one bf16 step)."""
import csv, io, json, os, subprocess, sys
rep, out_path, wl, keys = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    return float(v) * f
out = {}
for k, r in zip(keys, rows[2:2 + len(keys)]):
    rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
    wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    out[k] = rd + wr
    print(k, r[idx["Kernel Name"]][:60], rd + wr)
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
data[wl] = out
data["_source"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full ({os.path.basename(rep)}), see profiles/*.md"
json.dump(data, open(out_path, "w"), indent=1)
