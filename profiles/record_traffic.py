#!/usr/bin/env python
"""profiles/record_traffic.py WORKLOAD REPORT.ncu-rep -- record the DRAM bytes of one CG-kernel launch in
profiles/cg_traffic.json together with the kernel name and the hash of the sources the library was built from
(bench.kernel_source_hash()).  bench.py quotes `roofline.traffic` only when that hash and kernel match the library
it is timing, so a stale capture can never be passed off as a measurement of the current kernels."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    workload, rep = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, row = rows[0], rows[1], rows[2]

    def metric(name):
        i = names.index(name)
        v = float(row[i].replace(",", ""))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[i]]
        return v * scale

    rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
    kernel = row[names.index("Kernel Name")]
    path = os.path.join(ROOT, "profiles", "cg_traffic.json")
    try:
        data = json.load(open(path))
    except Exception:
        data = {}
    data["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the CG kernel (ncu --set full); "
                        "written by profiles/record_traffic.py; source_hash = bench.kernel_source_hash() of the tree "
                        "the captured library was built from")
    data[workload] = {"kernel": kernel.replace("void ", "").split("(")[0], "dram_bytes_per_launch": rd + wr,
                      "dram_bytes_read": rd, "dram_bytes_write": wr, "source_hash": bench.kernel_source_hash(),
                      "report": os.path.basename(rep)}
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(data[workload]))


if __name__ == "__main__":
    main()
