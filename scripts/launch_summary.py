"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
    python scripts/launch_summary.py gpurun_out/launches.csv out.csv "<comment>" [kernel-name regex]"""
import collections
import csv
import re
import sys


def main():
    src, dst, comment = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    keep = re.compile(sys.argv[4]) if len(sys.argv) > 4 else None
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        if keep is not None and not keep.search(name):
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# %s\n" % comment)
        f.write("kernel,launches_captured,avg_us,total_us,share_of_captured_kernel_time\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.2f,%.1f,%.4f\n" % (k, n, t / n, t, t / tot))
    print(open(dst).read())


if __name__ == "__main__":
    main()
