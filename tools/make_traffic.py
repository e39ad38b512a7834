"""profiles/roofline_traffic.json from an ncu --set full report: DRAM bytes (read + write) per launch, averaged per kernel.
    python tools/make_traffic.py gpurun_out/r01_full.ncu-rep profiles/roofline_traffic.json"""
import collections
import csv
import json
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.OrderedDict()
for r in data:
    name = re.sub(r"[<(].*", "", r[ix["Kernel Name"]]).split("::")[-1]
    b = sum(float(r[ix[k]].replace(",", "")) * scale[units[ix[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    t = float(r[ix["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[ix["gpu__time_duration.sum"]]]
    a = agg.setdefault(name, dict(launches=0, dram_bytes=0.0, us=0.0))
    a["launches"] += 1
    a["dram_bytes"] += b
    a["us"] += t
out = {k: dict(launches=v["launches"], dram_bytes_per_launch=v["dram_bytes"] / v["launches"], us_per_launch=v["us"] / v["launches"]) for k, v in agg.items()}
out["_source"] = "ncu --set full --clock-control none, first 16 launches of tools/prof_run.py 61440 (one 61440-row batch): " + sys.argv[1]
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
