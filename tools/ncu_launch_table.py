"""Print the mpn:: kernels of an ncu `--metrics gpu__time_duration.sum --csv` launch list (last `n` launches), one per line."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
out = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[1:] if 'mpn::' in r[ki]]
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(out)
tot = 0.0
for name, ns in out[-n:]:
    tot += ns
    print("%9.1f us  %s" % (ns / 1e3, name[:110]))
print("%9.1f us  total of %d launches" % (tot / 1e3, min(n, len(out))))
