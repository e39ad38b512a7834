"""Turn ncu exports into the small text summaries committed under profiles/.

    python tools/summarize_profile.py launches <launches.csv> <out.md>     (ncu --metrics gpu__time_duration.sum --csv)
    python tools/summarize_profile.py full <report.ncu-rep> <out.md>        (ncu --set full)
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).split("::")[-1]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.3f | %.1f |\n" % (k, v[0], v[1], v[1] / tot, v[1] / v[0]))
        f.write("| **total** | %d | %.1f | 1.000 | |\n" % (sum(v[0] for v in agg.values()), tot))


KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("| # | kernel | " + " | ".join(k.split(".")[0] for k in KEYS) + " |\n")
        f.write("|---|---|" + "---:|" * len(KEYS) + "\n")
        f.write("| | unit | " + " | ".join(units[idx[k]] if k in idx else "" for k in KEYS) + " |\n")
        for n, r in enumerate(data):
            name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).split("::")[-1]
            f.write("| %d | `%s` | " % (n, name) + " | ".join(r[idx[k]] if k in idx else "" for k in KEYS) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
