#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libppt_b200.so (cuobjdump -sass): the evidence that the tensor path is
tcgen05 (UTCHMMA / UTCQMMA), accumulators live in tensor memory (LDTM), operands move by bulk async copies
(UBLKCP) under mbarriers (SYNCS / UTCBAR), argmax uses CREDUX, and the bit-exact geometry kernels hold no FFMA
beyond the reference's own dot-product chain.

    python tools/sass_counts.py > profiles/sass_counts_r2.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ppt_b200", "libppt_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "SYNCS", "CREDUX", "REDUX", "FFMA", "FFMA2",
         "FADD", "FMUL", "HMMA", "ATOMG", "RED", "LDGSTS")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# %s: code objects %s, %d kernels" % (os.path.basename(LIB), ", ".join(archs), len(kernels)))
    tot = collections.Counter()
    for (name, c), pretty in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", pretty.replace("(anonymous namespace)::", ""))
        print("%-78s %6d instr  %s" % (short[:78], c["_total"], " ".join("%s=%d" % (w, c[w]) for w in WATCH if c[w])))
        tot.update(c)
    print("# total: " + " ".join("%s=%d" % (w, tot[w]) for w in WATCH if tot[w]))


if __name__ == "__main__":
    sys.exit(main())
