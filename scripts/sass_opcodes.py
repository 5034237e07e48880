"""cuobjdump -sass opcode histogram per object file of libnmx (the Blackwell-native evidence: UTCHMMA = tcgen05.mma,
UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy,
SYNCS = mbarrier ops, REDG = vector red.global).  Writes profiles/r2_sass_opcodes.txt."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "REDG", "ATOMG", "RED", "ELECT",
       "FADD2", "FFMA2", "F2FP", "MUFU", "HMMA", "LDSM", "STS", "LDS", "LDG", "STG", "MEMBAR", "FENCE", "CCTL"]
out = ["SASS opcode counts per object file (cuobjdump -sass of nerf_meets_mlx_b200/_lib/*.o, sm_100a).",
       "UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG/UTMASTG = cp.async.bulk.tensor load/store,",
       "UBLKCP = cp.async.bulk, SYNCS = mbarrier, REDG = red.global (vector reductions of the weight-gradient flush).",
       "HMMA / LDSM = mma.sync / ldmatrix: only in nmx_tiny.o, the byte-bound width-64 MLP (DESIGN.md 3.2).", ""]
for o in sorted(glob.glob(os.path.join(ROOT, "nerf_meets_mlx_b200", "_lib", "*.o"))):
    txt = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    ops = collections.Counter()
    kernels = re.findall(r"Function : (\S+)", txt)
    for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", txt, re.M):
        ops[m.group(1)] += 1
    line = ", ".join(f"{k} {ops[k]}" for k in KEY if ops.get(k))
    out.append(f"{os.path.basename(o)}: {len(kernels)} kernels, {sum(ops.values())} instructions")
    out.append(f"    {line}")
with open(os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt"), "w") as f:
    f.write("\n".join(out) + "\n")
print("\n".join(out))
