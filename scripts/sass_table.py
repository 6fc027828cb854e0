"""Per-kernel SASS evidence from the built libsic.so: registers, shared memory and counts of the mnemonics that prove which hardware
paths a kernel uses (tcgen05.mma = UTCHMMA, tcgen05.ld / st = LDTM / STTM, tcgen05.commit / mbarrier = UTCBAR / SYNCS, bulk L2
prefetch = UBLKPF, TMA tensor copies = UTMALDG / UTMASTG, special-function unit = MUFU, 128-bit global accesses = LDG.E.128 / STG.E.128).

    python scripts/sass_table.py [path/to/libsic.so] > profiles/<round>_sass_table.txt

Runs without a GPU (cuobjdump only reads the fatbin)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "domain_specific_image_compression_b200", "libsic.so")
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "SYNCS", "UBLKPF", "UTMALDG", "UTMASTG", "MUFU", "LDG.E.128", "STG.E.128", "LDS", "STS",
             "SHFL", "DADD", "ATOM", "RED"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"sic::", "", name)
    return re.sub(r"\(.*", "", name)


res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in line:
        f = dict(kv.split(":") for kv in line.split() if ":" in kv)
        usage[cur] = (int(f.get("REG", 0)), int(f.get("SHARED", 0)), int(f.get("LOCAL", 0)))
        cur = None

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + ".") or (("." in mn) and op.startswith(mn)):
                counts[cur][mn] += 1

names = demangle(list(counts))
print(f"# {os.path.relpath(lib, ROOT)}: cuobjdump -sass / -res-usage, sm_100a; one row per kernel, zero columns omitted")
print(f"# {'kernel':<70} {'regs':>4} {'smem':>7} {'local':>5} {'instr':>6}  mnemonic counts")
for k, c in counts.items():
    regs, smem, local = usage.get(k, (0, 0, 0))
    cols = " ".join(f"{mn}={c[mn]}" for mn in MNEMONICS if c[mn])
    print(f"{short(names.get(k, k))[:70]:<72} {regs:>4} {smem:>7} {local:>5} {c['_total']:>6}  {cols}")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("# library totals: " + " ".join(f"{mn}={tot[mn]}" for mn in MNEMONICS if tot[mn]))
