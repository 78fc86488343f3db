"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total time,
count and share.  usage: python scripts/summarize_launches.py launches.csv [steps] > summary.txt"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"<.*", "", name)[:90]
        try:
            t = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].strip(), 1e-6)
        tot[name] += t * scale
        cnt[name] += 1
    s = sum(tot.values())
    print(f"# {sum(cnt.values())} launches, {s:.2f} ms total device time over {steps} step(s) "
          f"= {s / steps:.2f} ms/step (ncu-serialised, cold caches: compare shares, not absolutes)")
    print("# share   ms_total   ms/step   launches  kernel")
    for n, t in tot.most_common(40):
        print(f"{t / s * 100:6.2f}%  {t:9.2f}  {t / steps:8.2f}  {cnt[n]:8d}  {n}")


if __name__ == "__main__":
    main()
