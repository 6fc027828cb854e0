"""Per-kernel instruction mix from the SASS page of an .ncu-rep (ncu --set full --import-source on): executed warp- and
thread-level counts of MUFU (the XU / SFU pipe), FP32 (FFMA/FMUL/FADD), tensor (UTC*MMA) and memory instructions, and the
achieved rates against the SM peaks.  This is how the cdf_diff kernel's "SFU-bound" claim is checked: achieved MUFU thread-ops/s
vs 148 SMs x 16 lanes x SM clock.
    python scripts/ncu_pipes.py gpurun_out/x.ncu-rep [--json out.json] [--clock-mhz 1965]"""
import csv, io, json, subprocess, sys

rep = sys.argv[1]
out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
clock = float(sys.argv[sys.argv.index("--clock-mhz") + 1]) * 1e6 if "--clock-mhz" in sys.argv else 1965e6
SMS = 148
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rrows = list(csv.reader(io.StringIO(raw)))
ridx = {h: i for i, h in enumerate(rrows[0])}
durs = []
for r in rrows[2:]:
    t = float(r[ridx["gpu__time_duration.sum"]]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}[rrows[1][ridx["gpu__time_duration.sum"]]]
    durs.append((r[ridx["Kernel Name"]], t, int(float(r[ridx["launch__registers_per_thread"]]))))
kernels, cur = [], None
for row in csv.reader(io.StringIO(src)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        if kernels and kernels[-1]["name"] == row[1] and not kernels[-1].get("dup"):
            cur = {"name": row[1], "cols": None, "ops": {}, "dup": True}     # the page prints most kernels twice (two views, same SASS)
        else:
            cur = {"name": row[1], "cols": None, "ops": {}}
            kernels.append(cur)
    elif row[0] == "Address":
        cur["cols"] = {h: i for i, h in enumerate(row)}
    elif cur is not None and cur["cols"] is not None and row[0].startswith("0x"):
        c = cur["cols"]
        sass = row[c["Source"]].split()
        if not sass:
            continue
        op = sass[1] if sass[0].startswith("@") and len(sass) > 1 else sass[0]
        base = op.split(".")[0]
        w, t = int(row[c["Instructions Executed"]]), int(row[c["Thread Instructions Executed"]])
        a = cur["ops"].setdefault(base, [0, 0]); a[0] += w; a[1] += t
        if base == "MUFU":
            a = cur["ops"].setdefault(op, [0, 0]); a[0] += w; a[1] += t
res = []
for k, (nm, t, regs) in zip(kernels, durs):
    ops = k["ops"]
    tot_w = sum(v[0] for kk, v in ops.items() if "." not in kk)
    mufu_t = ops.get("MUFU", [0, 0])[1]
    fp32_w = sum(ops.get(o, [0, 0])[0] for o in ("FFMA", "FMUL", "FADD", "FSEL", "FSETP", "FMNMX", "FCHK"))
    peak_mufu = SMS * 16 * clock
    peak_issue = SMS * 4 * clock
    e = {"kernel": nm[:110], "registers": regs, "duration_us": t * 1e6, "warp_inst": tot_w, "mufu_thread_ops": mufu_t,
         "mufu_breakdown_warp_inst": {kk: v[0] for kk, v in ops.items() if kk.startswith("MUFU.")},
         "mufu_ops_per_s": mufu_t / t, "mufu_peak_ops_per_s": peak_mufu, "mufu_frac_of_peak": mufu_t / t / peak_mufu,
         "issue_frac_of_peak": tot_w / t / peak_issue, "fp32_warp_inst": fp32_w,
         "top_ops": sorted(((kk, v[0]) for kk, v in ops.items() if "." not in kk), key=lambda z: -z[1])[:10]}
    res.append(e)
    print(f"{e['kernel']}\n    {regs} regs, {e['duration_us']:.1f} us, {tot_w/1e6:.1f} M warp instr (issue {100*e['issue_frac_of_peak']:.1f}% of 4/clk/SM), "
          f"MUFU {mufu_t/1e9:.3f} G thread-ops -> {e['mufu_ops_per_s']/1e12:.2f} T/s = {100*e['mufu_frac_of_peak']:.1f}% of {peak_mufu/1e12:.2f} T/s "
          f"(148 SMs x 16 lanes x {clock/1e6:.0f} MHz)\n    MUFU mix: {e['mufu_breakdown_warp_inst']}\n    top: {e['top_ops']}")
if out_json:
    json.dump(res, open(out_json, "w"), indent=1)
