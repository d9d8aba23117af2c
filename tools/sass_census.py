"""SASS census of libmpn_b200.so: per kernel, the counts of the mnemonics that prove which hardware paths are used
(tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM, TMA = UTMALDG, tcgen05.commit / mbarrier = UTCBAR / SYNCS, cp.async = LDGSTS,
packed fp32 = FFMA2 / FADD2, fp64 = DFMA / DADD).  Runs without a GPU (cuobjdump).

    python tools/sass_census.py [out.md]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "graph-convolutional-network-for-multi-camera-vehicle-tracking_b200", "libmpn_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "SYNCS", "LDGSTS", "FFMA2", "FADD2",
         "DFMA", "DADD", "MUFU", "ATOMG", "REDG", "ACQBULK", "BAR"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur]["_all"] += 1
            op = m.group(1)
            for w in WATCH:
                if op.startswith(w):
                    kernels[cur][w] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    lines = ["# SASS census of libmpn_b200.so (cuobjdump -sass, sm_100a)", "",
             "Columns: instructions in the kernel, then the watched mnemonics that occur (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, "
             "UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGSTS = cp.async, FFMA2 / FADD2 = packed fp32).", "",
             "| kernel | SASS instructions | tensor / async mnemonics |", "|---|---:|---|"]
    tot = collections.Counter()
    for (mangled, c), name in zip(kernels.items(), demangled):
        short = re.sub(r"\(.*", "", name)[:90]
        hits = ", ".join("%s %d" % (w, c[w]) for w in WATCH if c[w])
        lines.append("| `%s` | %d | %s |" % (short, c["_all"], hits))
        tot.update(c)
    lines += ["", "Totals: %d kernels, %d instructions; %s" % (len(kernels), tot["_all"], ", ".join("%s %d" % (w, tot[w]) for w in WATCH if tot[w]))]
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text[-1500:])


if __name__ == "__main__":
    main()
