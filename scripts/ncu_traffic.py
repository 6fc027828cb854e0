"""Append the per-launch DRAM traffic of every kernel in an .ncu-rep (ncu --set full) to profiles/ncu_traffic.json.
bench.py quotes `roofline.traffic` from that file, and only when the capture's register count equals the loaded library's
(sic_kernel_registers), so the figure always describes the binary that ran.
    python scripts/ncu_traffic.py gpurun_out/x.ncu-rep --shape 16,128,256,256 [--match gdn_] [--note "..."]"""
import csv, io, json, os, re, subprocess, sys
rep = sys.argv[1]
arg = lambda k, d=None: sys.argv[sys.argv.index(k) + 1] if k in sys.argv else d
shape = [int(v) for v in arg("--shape", "").split(",") if v]
match, note = arg("--match", ""), arg("--note", "")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
idx, units = {h: i for i, h in enumerate(rows[0])}, rows[1]
def val(r, k):
    u = units[idx[k]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return float(r[idx[k]]) * scale
db = json.load(open(path)) if os.path.exists(path) else []
for r in rows[2:]:
    m = re.search(r"(\w+)<([^>]*)>\(", r[idx["Kernel Name"]]) or re.search(r"(\w+)()\(", r[idx["Kernel Name"]])
    name = m.group(1) + ("<" + re.sub(r"\(\w+\)|\s", "", m.group(2)) + ">" if m.group(2) else "")
    if match and match not in name:
        continue
    e = {"kernel": name, "shape": shape, "registers": int(val(r, "launch__registers_per_thread")), "grid": int(val(r, "launch__grid_size")),
         "block": int(val(r, "launch__block_size")), "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
         "duration_us_under_ncu": val(r, "gpu__time_duration.sum") * 1e6, "capture": os.path.basename(rep), "note": note}
    db = [d for d in db if not (d["kernel"] == name and d["shape"] == shape and d["registers"] == e["registers"])] + [e]
    print(e)
json.dump(db, open(path, "w"), indent=1)
