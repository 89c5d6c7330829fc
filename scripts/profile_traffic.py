"""profiles/r01_dram_traffic.json: DRAM bytes (read+write) per launch of the three big kernels, from an
`ncu --set full` report of `bench.py --workload W --steps 1` (launch order per step: contract fwd, dphi, contract bwd).
usage: python scripts/profile_traffic.py REPORT.ncu-rep WORKLOAD"""
import csv, io, json, os, subprocess, sys
rep, wl = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    return float(v) * f
keys = ["qc_contract.fwd", "qc_dphi.bwd", "qc_contract.bwd"]
out = {}
for k, r in zip(keys, rows[2:5]):
    rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
    wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    out[k] = rd + wr
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r01_dram_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[wl] = out
data["_source"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, see profiles/*.md"
json.dump(data, open(path, "w"), indent=1)
print(data)
