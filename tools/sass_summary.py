"""Per-kernel SASS evidence for the built library: counts of the instructions that tell which
hardware path a kernel uses (DMMA = FP64 tensor pipe, DFMA/DADD/DMUL = FP64 ALU, LDGSTS = cp.async,
UBLKCP = bulk copy (cp.async.bulk, the 1-D TMA path), UTMALDG = tensor-map TMA, UTC*MMA / LDTM = tcgen05 / TMEM), registers and spills from cuobjdump -res-usage.

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "alabi_b200", "libalabi_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
usage = {}
cur = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", ln)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
keys = ["DMMA", "DFMA", "DADD", "DMUL", "LDGSTS", "UBLKCP", "UTMALDG", "UTCMMA", "LDTM", "LDS", "LDG", "BAR"]
counts, arch = collections.OrderedDict(), set()
cur = None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1)
        for k in keys:
            if op.startswith(k) or (k == "UTCMMA" and op.startswith("UTC") and "MMA" in op):
                counts[cur][k] += 1
print(f"# cuobjdump -sass / -res-usage of alabi_b200/libalabi_b200.so; cubin architectures: {sorted(arch)}")
print(f"# {len(counts)} kernels; totals: " + ", ".join(f"{k} {sum(c[k] for c in counts.values())}" for k in keys))
print(f"{'kernel':70s} {'regs':>5s} {'smem':>7s} {'local':>6s} " + " ".join(f"{k:>7s}" for k in keys))
for name, c in sorted(counts.items(), key=lambda kv: -(kv[1]['DMMA'] * 1000 + kv[1]['DFMA'])):
    short = re.sub(r"\(anonymous namespace\)::", "", demangle(name)).split("(")[0][:70]
    r, sh, lo = usage.get(name, (0, 0, 0))
    print(f"{short:70s} {r:5d} {sh:7d} {lo:6d} " + " ".join(f"{c[k]:7d}" for k in keys))
