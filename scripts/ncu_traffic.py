"""Extract per-launch DRAM traffic from `ncu --set full` reports into profiles/<tag>_traffic.json.

usage: python scripts/ncu_traffic.py <tag> "<key>=<report.ncu-rep>" ...
The key is the one bench.py looks up: "<workload> B=<batch> <step kind> <fwd|bwd> F=<folds>".
"""
import csv, io, json, os, subprocess, sys

tag = sys.argv[1]
out = {}
for arg in sys.argv[2:]:
    key, rep = arg.rsplit("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def num(v):
        try:
            return float(v.replace(",", ""))
        except ValueError:
            return v

    m = {h: (num(v), u) for h, u, v in zip(hdr, units, vals)}

    def to_bytes(name):
        v, u = m[name]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return v * scale

    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    dur, du = m["gpu__time_duration.sum"]
    out[key] = {
        "kernel": m["Kernel Name"][0] if "Kernel Name" in m else "",
        "dram_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
        "duration_under_ncu": f"{dur} {du}", "report": os.path.basename(rep),
    }
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"{tag}_traffic.json")
with open(path, "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out, indent=1))
