#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of `bench.py` (the four main
kernels of one device-resident step on the bench workload): DRAM bytes and executed warp
instructions per step (summed over the launches of a kernel), stamped with the hash of the kernel sources so that bench.py only
reports them while they describe the kernels that are running.

    python tools/make_traffic.py gpurun_out/rNN_bench_full.ncu-rep > profiles/traffic.json
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    name_i = hdr.index("Kernel Name")
    rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    inst = hdr.index("smsp__inst_executed.sum")
    units = rows[1]

    def to_bytes(v, unit):
        v = float(v.replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return int(v * scale)

    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum per "
                       "step (summed over the launches of a kernel), ncu --set full on `bench.py --steps 1 --warmup 3` (workload configs[1]); "
                       "bench.py reports them only while csrc_sha16 matches the kernel sources",
           "source": Path(rep).name, "workload_bytes": bench.WORKLOAD["n"],
           "csrc_sha16": bench.csrc_sha16()}
    for r in rows[2:]:
        k = r[name_i].split("(")[0].split("<")[0].replace("void ", "").replace("lz77::", "")
        # (summed over the launches of the step: the encoder runs its input in pieces)
        out[k] = out.get(k, 0) + to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
        out[k + ".inst_executed"] = out.get(k + ".inst_executed", 0) + int(float(r[inst].replace(",", "")))
        out[k + ".launches"] = out.get(k + ".launches", 0) + 1
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
