"""Cuts the timed steps out of a full `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py.
A step starts at a k_integrate launch; the timed steps are numbers [preroll + warmup, preroll + warmup + K).
usage: ncu_cut_steps.py full.csv preroll warmup K out.csv   (prints the per-kernel share table)"""
import csv, sys
from collections import OrderedDict

def main(path, preroll, warmup, k, out):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    hdr = None
    for r in csv.reader(lines):
        if r[0] == "ID":
            hdr = r
            continue
        name = r[hdr.index("Kernel Name")]
        val = float(r[hdr.index("Metric Value")].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        rows.append((name, val / 1000.0 if unit == "ns" else val))
    starts = [i for i, (n, _) in enumerate(rows) if n.startswith("k_integrate")]
    a = starts[preroll + warmup]
    b = starts[preroll + warmup + k]
    cut, keep = [], True
    for n, t in rows[a:b]:                     # a step ends at k_head: drop statistics kernels launched between steps
        if n.startswith("k_integrate"):
            keep = True
        if keep:
            cut.append((n, t))
        if n.startswith("k_head"):
            keep = False
    with open(out, "w") as f:
        f.write('"ID","Kernel Name","gpu__time_duration.sum [us]"\n')
        for i, (n, t) in enumerate(cut):
            f.write('%d,"%s",%.3f\n' % (i, n, t))
    tot = sum(t for _, t in cut)
    agg = OrderedDict()
    for n, t in cut:
        key = n.split("(")[0]
        c, s = agg.get(key, (0, 0.0))
        agg[key] = (c + 1, s + t)
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for n, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (n, c, s, 100 * s / tot))
    print("total %.1f us over %d steps (%d launches)" % (tot, k, len(cut)))

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
