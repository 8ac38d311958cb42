"""Summaries of ncu output for profiles/.

  python tools/ncu_summary.py launches <launch_list.csv>
      per-kernel totals of a `ncu --metrics gpu__time_duration.sum --csv --log-file ...` launch list
  python tools/ncu_summary.py full <report.ncu-rep | raw.csv>
      key metrics and the main stall reasons of every kernel in a `--set full` capture
"""
import collections
import csv
import subprocess
import sys

KEY = [
  "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
  "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
  "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
  "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
  "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
  "smsp__warps_eligible.avg.per_cycle_active",
]


def short(name):
  name = name.split("(")[0]
  for junk in ("void ", "ign::"):
    name = name.replace(junk, "")
  return name[:70]


def launches(path):
  rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
  hdr = rows[0]
  ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
  tot, cnt = collections.Counter(), collections.Counter()
  for r in rows[1:]:
    if r[mi] == "gpu__time_duration.sum":
      tot[short(r[ki])] += float(r[vi].replace(",", "")) / 1e6
      cnt[short(r[ki])] += 1
  total = sum(tot.values())
  print("%-70s %8s %10s %6s" % ("kernel", "launches", "total ms", "share"))
  for k, ms in tot.most_common():
    print("%-70s %8d %10.3f %5.1f%%" % (k, cnt[k], ms, 100 * ms / total))
  print("%-70s %8d %10.3f" % ("all", sum(cnt.values()), total))


def full(path):
  if path.endswith(".ncu-rep"):
    text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
  else:
    rows = list(csv.reader(open(path, errors="replace")))
  hdr, units = rows[0], rows[1]
  ki = hdr.index("Kernel Name")
  for vals in rows[2:]:
    if len(vals) != len(hdr):
      continue
    print("==", short(vals[ki]))
    for k in KEY:
      if k in hdr:
        print("  %-58s %s %s" % (k, vals[hdr.index(k)], units[hdr.index(k)]))
    stalls = []
    for i, h in enumerate(hdr):
      if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        try:
          stalls.append((float(vals[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        except ValueError:
          pass
    print("  stalls per issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:6]))


if __name__ == "__main__":
  if len(sys.argv) != 3 or sys.argv[1] not in ("launches", "full"):
    sys.exit(__doc__)
  (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
