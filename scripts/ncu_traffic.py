"""DRAM / L2 bytes and tensor-pipe share of one kernel from an `ncu --set full` report -> profiles/dominant_kernel_traffic.json
(read by bench.py at run time for roofline.traffic).
usage: python scripts/ncu_traffic.py gpurun_out/prof.ncu-rep <kernel regex> profiles/dominant_kernel_traffic.json"""
import csv
import json
import re
import subprocess
import sys

rep, pat, out = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, k):
    v, u = float(r[ix[k]].replace(",", "")), units[ix[k]]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)


sel = [r for r in data if re.search(pat, r[ix["Kernel Name"]])]
if not sel:
    raise SystemExit("no kernel matching %r" % pat)
n = len(sel)
d = {"kernel": re.sub(r"\(.*", "", sel[0][ix["Kernel Name"]]).replace("void ", "").replace("fgc::", ""),
     "launches_averaged": n,
     "dram_bytes_per_launch": sum(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in sel) / n,
     "l2_to_sm_bytes_per_launch": sum(val(r, "l1tex__m_xbar2l1tex_read_bytes.sum") for r in sel) / n,
     "tensor_pipe_pct": sum(float(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]) for r in sel) / n,
     "duration_us": sum(val(r, "gpu__time_duration.sum") for r in sel) / n / (1e3 if units[ix["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1.0),
     "source": "ncu --set full --clock-control none, %s" % rep.split("/")[-1]}
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d))
