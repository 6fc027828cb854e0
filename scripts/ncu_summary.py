"""Condense an .ncu-rep (ncu --set full) into the handful of metrics the roofline argument uses.
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
M = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
     "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
     "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
print(f"# {rep}: ncu --set full --clock-control none (cold-cache, serialised replays: compare shares, not absolutes)")
for r in data:
    print("\n" + r[idx["Kernel Name"]][:150])
    for m in M:
        if m in idx:
            print(f"    {m:95s} {r[idx[m]]:>18s} {units[idx[m]]}")
    try:
        t = float(r[idx["gpu__time_duration.sum"]]); tu = units[idx["gpu__time_duration.sum"]]
        t *= {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[tu]
        def gb(k):
            v = float(r[idx[k]]); u = units[idx[k]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]
        tr = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        print(f"    {'=> DRAM traffic per launch / duration':95s} {tr/1e6:14.1f} MB  {tr/t/1e9:8.1f} GB/s")
    except Exception as e:
        pass
