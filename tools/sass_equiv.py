"""Device-code comparison of two object files (or .so), kernel by kernel: identical SASS / same after masking uniform-register
numbers (ptxas picks among equivalent allocations) / different.  Used to show that a change that is compiled out or switched
off by default leaves the shipped kernels untouched when no GPU is at hand:

    git stash; bash csrc/build.sh; cp -r csrc/_obj /tmp/before; git stash pop; bash csrc/build.sh
    python tools/sass_equiv.py /tmp/before/gemm_tc.o csrc/_obj/gemm_tc.o
"""
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    table, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            table[name] = []
        elif name and "/*" in line:
            ins = " ".join(re.sub(r"/\*[0-9a-fx ]+\*/", "", line).split())       # drop addresses and encodings
            if ins:
                table[name].append(ins)
    return table


def main():
    a, b = functions(sys.argv[1]), functions(sys.argv[2])
    worst = 0
    for name in sorted(set(a) | set(b)):
        if name not in a or name not in b:
            print("%-12s %s" % ("only in " + ("first" if name in a else "second"), name))
            continue
        if a[name] == b[name]:
            verdict = "identical"
        elif [re.sub(r"UR\d+", "UR", x) for x in a[name]] == [re.sub(r"UR\d+", "UR", x) for x in b[name]]:
            verdict = "UR-renamed"
        else:
            verdict, worst = "DIFFERENT", 1
        print("%-12s %5d %s" % (verdict, len(b[name]), name))
    sys.exit(worst)


if __name__ == "__main__":
    main()
