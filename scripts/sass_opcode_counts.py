"""Regenerates profiles/r02_sass_opcode_counts.txt: per-kernel SASS opcode counts of the in-tree library (runs without a GPU).
    python scripts/sass_opcode_counts.py"""
import collections, os, re, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "e2e_tts_b200", "lib", "libe2e_tts_b200.so")
KEEP = ["UTCHMMA.2CTA", "UTCHMMA", "HMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "UTCBAR", "SYNCS", "FADD2", "FMUL2",
        "FFMA2", "MUFU"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names, counts, total = [], {}, {}
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        names.append(cur)
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        base = op.split(".")[0]
        if base == "UTCHMMA" and ".2CTA" in op:
            counts[cur]["UTCHMMA.2CTA"] += 1
        elif base in KEEP:
            counts[cur][base] += 1
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
out = ["cuobjdump -sass e2e_tts_b200/lib/libe2e_tts_b200.so (built by e2e_tts_b200/build.py: nvcc -gencode arch=compute_100a,code=sm_100a "
       "-lineinfo -O3); regenerate with scripts/sass_opcode_counts.py",
       "opcode counts per kernel (SASS mnemonics: UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA "
       "tensor load/store,",
       "UBLKCP = cp.async.bulk, UBLKPF = cp.async.bulk.prefetch.L2, UTCBAR = tcgen05.commit, FADD2/FMUL2/FFMA2 = packed f32x2; HMMA (legacy mma.sync) must be 0)", ""]
hmma = 0
for n, d in zip(names, dem):
    d = d.replace("(anonymous namespace)::", "").replace("(bool)", "")
    short = re.sub(r"^void |e2e::", "", d.split("(")[0]).replace(", ", ",")
    c = counts[n]
    hmma += c["HMMA"]
    out.append("%-36s %5d instr  %s" % (short[:36], total[n], " ".join("%s=%d" % (k, c[k]) for k in KEEP if c[k])))
out.append("")
out.append("HMMA total: %d" % hmma)
open(os.path.join(ROOT, "profiles", "r02_sass_opcode_counts.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-8:]))
