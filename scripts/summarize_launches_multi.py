"""Summarise an ncu launch list with several metrics per launch (gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum): per kernel name launches, total / mean time, mean DRAM bytes per launch.
usage: python scripts/summarize_launches_multi.py launches.csv > summary.txt"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, ni, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    t, rd, wr, cnt = (collections.Counter() for _ in range(4))
    seen = set()
    unit = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r[ki]))[:80]
        try:
            v = float(r[vi].replace(",", "")) * unit.get(r[ui].strip(), 1.0)
        except ValueError:
            continue
        if (r[ii], name) not in seen:
            seen.add((r[ii], name))
            cnt[name] += 1
        if r[ni].startswith("gpu__time_duration"):
            t[name] += v
        elif r[ni].startswith("dram__bytes_read"):
            rd[name] += v
        elif r[ni].startswith("dram__bytes_write"):
            wr[name] += v
    s = sum(t.values())
    print(f"# {sum(cnt.values())} launches, {s:.2f} ms total device time (ncu-serialised, cold caches: compare shares, "
          f"not absolutes); DRAM bytes are per-launch means")
    print("# share   ms_total  launches  ms/launch  dram_read_MB/launch  dram_write_MB/launch  kernel")
    for n, tt in t.most_common(45):
        c = cnt[n]
        print(f"{tt / s * 100:6.2f}%  {tt:9.2f}  {c:8d}  {tt / c:9.4f}  {rd[n] / c / 1e6:12.2f}  {wr[n] / c / 1e6:12.2f}  {n}")


if __name__ == "__main__":
    main()
