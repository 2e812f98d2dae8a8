"""SASS opcode histogram of the built library (cuobjdump -sass): the Blackwell-native evidence kept under profiles/.
    python scripts/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ray3d_b200", "libray3d_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], check=True, capture_output=True, text=True).stdout
ops, variants, per_fn = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
fn = "?"
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m:
        ops[m.group(1)] += 1
        variants[m.group(1) + m.group(2)] += 1
        per_fn[fn][m.group(1)] += 1
print("# SASS opcode histogram of ray3d_b200/libray3d_b200.so (cuobjdump -sass, sm_100a), round 2, default build")
print("# Blackwell-native evidence: UTCHMMA = tcgen05.mma (kind::f16), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor loads/stores,")
print("# UTCBAR = tcgen05.commit, FADD2/FMUL2/FFMA2 = packed fp32x2 ALU ops of sm_100; no HMMA (legacy mma.sync)")
for k in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCCP", "UBLKCP", "SYNCS", "FADD2", "FMUL2", "FFMA2", "HMMA", "IMMA",
          "DFMA", "DMUL", "DADD", "REDG", "ATOMG", "LDG", "STG", "LDS", "STS", "FFMA"):
    print(f"{k:10s} {ops.get(k, 0):7d}")
print("# variants")
for k, v in sorted(variants.items()):
    if k.split(".")[0] in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "REDG", "ATOMG"):
        print(f"{k:44s} {v:5d}")
print("# per kernel: tcgen05.mma / tcgen05.ld / TMA load / TMA store / packed fp32x2")
for f, c in sorted(per_fn.items()):
    if c.get("UTCHMMA") or c.get("UTMALDG"):
        print(f"{f[:100]:100s} {c.get('UTCHMMA', 0):4d} {c.get('LDTM', 0):4d} {c.get('UTMALDG', 0):4d} {c.get('UTMASTG', 0):4d} {c.get('FADD2', 0) + c.get('FMUL2', 0) + c.get('FFMA2', 0):4d}")
