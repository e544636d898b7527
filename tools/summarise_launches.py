"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --steps S --warmup W` into per-kernel
totals for ONE training step (the last complete step of the capture).

usage: python tools/summarise_launches.py raw.csv n_steps_total > one_step.csv
       python tools/summarise_launches.py raw.csv --marker k_csr_hist > one_step.csv     (a step starts at every launch of the
                                                       marker kernel: the CSR build opens each training step; takes the last complete step)
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("gnnml3::", "")
    m = re.match(r"([A-Za-z_0-9]+(<[^(]*>)?)\(", name)
    if name.startswith("at::") or "at::" in name[:40]:
        return "torch: " + name[-60:].replace('"', "'")
    return m.group(1) if m else name[:60]


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
    if sys.argv[2] == "--marker":
        starts = [i for i, r in enumerate(rows) if sys.argv[3] in r[4]]
        # the capture may end in the middle of a step (-c N): take the last COMPLETE step when there is one
        last = rows[starts[-2]:starts[-1]] if len(starts) >= 2 else rows[starts[-1]:]
    else:
        steps = int(sys.argv[2])
        per = len(rows) // steps
        last = rows[len(rows) - per:]
    tot = collections.OrderedDict()
    for r in last:
        k = short(r[4])
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += float(r[-1].replace(",", "")) / 1e3
    total = sum(v[1] for v in tot.values())
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches_per_step", "total_us_per_step", "share_of_step_gpu_time"])
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, v[0], round(v[1], 1), round(v[1] / total, 4)])
    w.writerow(["TOTAL", sum(v[0] for v in tot.values()), round(total, 1), 1.0])


if __name__ == "__main__":
    main()
