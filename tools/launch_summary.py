"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python tools/launch_summary.py gpurun_out/launches.csv > profiles/launches_summary.txt"""
import collections
import csv
import re
import sys


def main(path, note=""):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[0] == "ID":
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    total = sum(a[1] for a in agg.values())
    print("kernel, launches, total_us, share  (%s; cold-cache, serialised: compare SHARES)" % note)
    for name, (count, ns) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%s, %d, %.1f, %.1f%%" % (name, count, ns / 1e3, 100 * ns / total))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
