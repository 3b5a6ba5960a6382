#!/usr/bin/env python3
"""SASS opcode histogram per kernel of libcacfe.so (runs without a GPU: cuobjdump -sass):

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt

Per kernel: instruction count, registers are in the ncu digests; here the opcode counts that prove what the kernel is made of --
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk, the 1-D TMA), SYNCS
(mbarrier), FADD2 / FMUL2 / FFMA2 (packed f32x2), DFMA / DADD / DMUL (FP64), MUFU, LDS / STS / LDG / STG, BAR, SHFL, LDL / STL
(spills) -- and the ten most frequent opcodes.
"""
import collections, os, re, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "audio-training_b200", "libcacfe.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "FADD2", "FMUL2", "FFMA2", "FFMA", "FADD", "FMUL",
         "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "LDL", "STL", "HMMA"]
parts = re.split(r"\n\s*Function : ", txt)[1:]
for part, name in zip(parts, names):
    ops = collections.Counter()
    for line in part.split("\n"):
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            ops[m.group(1)] += 1
    total = sum(ops.values())
    short = (name[:name.rindex(">(") + 1] if ">(" in name else re.sub(r"\(.*", "", name)).replace("cacfe::", "").replace("void ", "")
    short = short.replace("(int)", "").replace("(bool)", "")
    watched = "  ".join(f"{k}:{ops[k]}" for k in WATCH if ops.get(k))
    top = "  ".join(f"{k}:{v}" for k, v in ops.most_common(10))
    print(f"== {short}   [{total} SASS instructions]\n   watched: {watched}\n   top10:   {top}")
