#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` report: dram__bytes_read.sum + dram__bytes_write.sum per launch for the
kernels bench.py's roofline names (mean over the captured launches of each kernel).
    python tools/ncu_traffic.py gpurun_out/<tag>/prof.ncu-rep profiles/ncu_traffic.json"""
import csv
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
NAMES = {"k_dgemm_nt": "k_dgemm_nt", "k_soap_forward": "soap_forward", "k_soap_adjoint": "soap_adjoint", "k_neigh": "connect"}

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
acc = {}
for d in data:
    name = d[ix["Kernel Name"]]
    for key, label in NAMES.items():
        if key in name:
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(d[ix[m]].replace(",", "")) * UNIT[units[ix[m]]]
            acc.setdefault(label, []).append(tot)
res = {k: sum(v) / len(v) for k, v in acc.items()}
res["_source"] = "%s: mean dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes)" % sys.argv[1]
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(res)
