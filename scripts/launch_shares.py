"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel shares of the step."""
import collections, csv, re, sys
path, steps = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr, data = None, []
for r in rows:
    if r[0] == "ID":
        hdr = r
    elif hdr and r[0].isdigit():
        data.append(r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
for r in data:
    full = r[ik].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    ours = "sic::" in r[ik] or re.search(r"(bottleneck_(fwd|bwd)_kernel|gdn_(fwd|bwd|dense)|cdf_tables_kernel|minmax|symbols_kernel)", full)
    name = re.sub(r"<.*", "", re.sub(r"\(.*", "", full)).split("::")[-1].replace("void ", "")[:58]
    key = ("[sic] " if ours else "      ") + name
    v = float(r[iv].replace(",", "")) / (1000.0 if r[iu] == "ns" else 1.0)
    agg[key][0] += 1
    agg[key][1] += v
    tot += v
print(f"# {len(data)} launches in {steps} timed steps; {tot/steps/1000:.3f} ms of kernel time per step (cold-cache, serialised: compare SHARES)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"{k:66s} n={n/steps:6.1f}/step {t/steps:10.1f} us/step {100*t/tot:5.1f}%")
sic = sum(t for k, (n, t) in agg.items() if k.startswith("[sic]"))
nsic = sum(n for k, (n, t) in agg.items() if k.startswith("[sic]"))
print(f"# our kernels: {nsic/steps:.0f} launches/step, {sic/steps:.1f} us/step = {100*sic/tot:.1f}% of the step's kernel time")
