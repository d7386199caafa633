#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libblstm_b200.so (cuobjdump -sass): the Blackwell-native opcodes each kernel contains.
    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, SYNCS = mbarrier, DFMA / DMUL / DADD = the FP64
chain of the reference-exact exp, FFMA2 = packed fp32 FMA.  Runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "lstm-rnn_b200", "libblstm_b200.so")
text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
want = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "SYNCS", "ELECT", "DFMA", "DMUL", "DADD", "FFMA2", "FFMA", "MUFU", "HMMA", "LDG", "STG", "LDS", "STS", "BAR", "SHFL"]
kern, total = collections.OrderedDict(), collections.Counter()
name = None
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        base = op.split(".")[0]
        kern[name]["instructions"] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            kern[name]["UTCHMMA.2CTA"] += 1
        if base in want:
            kern[name][base] += 1
print("# cuobjdump -sass lstm-rnn_b200/libblstm_b200.so, opcode counts per kernel (static SASS, sm_100a)")
cols = ["instructions"] + want
print("%-72s" % "kernel" + "".join("%9s" % c[:9] for c in cols))
for k, c in kern.items():
    if c["instructions"] < 50:
        continue
    print("%-72s" % k[:72] + "".join("%9d" % c[x] for x in cols))
