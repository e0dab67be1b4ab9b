"""Per-kernel counts of the tensor-core / TMA / TMEM / cluster mnemonics in the in-tree libmidivae.so (cuobjdump -sass): the static proof that the
shipped kernels are tcgen05 / TMA / cluster code.  python scripts/sass_excerpt.py > profiles/r2/sass_excerpt.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "midi_vae_b200/libmidivae.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]+)")
keep = re.compile(r"^(UTCHMMA|UTMALDG|UTMAPF|UBLKCP|LDTM|STTM|UTCBAR|UTCATOMSWS|STSM|LDSM|HMMA|SYNCS|UCGABAR|REDG|ATOMG|ATOMS|CCTL|MEMBAR|REDUX|STAS)")
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    if "Function :" in line:
        kern = line.split("Function :")[1].strip()
        counts[kern] = collections.Counter()
        continue
    m = pat.match(line)
    if m and kern and keep.match(m.group(1)):
        counts[kern][m.group(1)] += 1
print("SASS evidence of the in-tree libmidivae.so (cuobjdump -sass, sm_100a), round 2 final build (scripts/sass_excerpt.py).")
print("Per kernel: counts of the tensor-core / TMA / TMEM / cluster mnemonics (B200_PROFILING.md: UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load,")
print("UBLKCP = cp.async.bulk, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, STSM / LDSM = stmatrix / ldmatrix, HMMA = mma.sync, SYNCS = mbarrier ops,")
print("REDG...F32x4 = red.global.add.v4.f32).  Only kernels that contain at least one of UTCHMMA / UTMALDG / UBLKCP / HMMA are listed.\n")
for k, c in counts.items():
    if not any(x.startswith(("UTCHMMA", "UTMALDG", "UBLKCP", "HMMA")) for x in c):
        continue
    short = re.sub(r"_ZN4mvae\d+_GLOBAL__N__[0-9a-f]+_\d+_(\w+?)_cu_[0-9a-f]+\d*", r"\1::", k)
    print(short)
    print("   " + ", ".join(f"{m} x{n}" for m, n in sorted(c.items())))
