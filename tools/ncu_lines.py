#!/usr/bin/env python
"""Per-source-line view of an `ncu --set full --import-source on` capture.

ncu's CSV source page lists SASS instructions with their executed counts and stall
samples; `nvdisasm -gi` lists the same SASS with the CUDA line each instruction came
from (the library is compiled with -lineinfo).  This joins the two by instruction
offset and prints, per source line of the kernel: warp instructions executed per unit
(token), share of all stall samples, average active threads.

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTRING CUBIN UNITS [--top N]

e.g. UNITS = tokens processed by the captured launch.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def sass_rows(report):
    raw = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    out, hdr = [], None
    kernel = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            kernel = r[1]
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) >= len(hdr) - 2 and r[0].startswith("0x"):
            out.append((kernel, dict(zip(hdr, r))))
    return out


def line_table(cubin, kernel_sub):
    """offset -> (file line, inlined-at chain text) for the first function whose name
    contains kernel_sub."""
    txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    table, cur, active = {}, None, False
    for line in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", line) or re.match(r"\s*//-+ \.text\.(\S+)", line)
        if m:
            active = kernel_sub in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3).strip())
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    report, ksub, cubin, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 60
    ksub_report = sys.argv[sys.argv.index("--name") + 1] if "--name" in sys.argv else ksub
    rows = [(k, r) for k, r in sass_rows(report) if ksub_report in (k or "")]
    if not rows:
        sys.exit("kernel not found in report")
    base = int(rows[0][1]["Address"], 16)
    table = line_table(cubin, ksub)
    agg = defaultdict(lambda: [0.0, 0.0, 0.0])  # inst, samples, thread-inst
    tot_inst = tot_samp = 0.0
    for _, r in rows:
        off = int(r["Address"], 16) - base
        inst = float(r["Instructions Executed"] or 0)
        samp = float(r["Warp Stall Sampling (All Samples)"] or 0)
        thr = float(r["Thread Instructions Executed"] or 0)
        f, ln, inl = table.get(off, ("?", 0, ""))
        key = (f, ln)
        agg[key][0] += inst
        agg[key][1] += samp
        agg[key][2] += thr
        tot_inst += inst
        tot_samp += samp
    print(f"{rows[0][0][:110]}")
    print(f"total warp instructions {tot_inst:.0f} = {tot_inst / units:.2f} per unit; "
          f"{tot_samp:.0f} stall samples")
    print(f"{'file:line':28s} {'inst/unit':>9s} {'inst %':>7s} {'stall %':>8s} {'thr':>5s}")
    for (f, ln), (inst, samp, thr) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f + ':' + str(ln):28s} {inst / units:9.2f} {100 * inst / tot_inst:6.1f}% "
              f"{100 * samp / max(tot_samp, 1):7.1f}% {thr / max(inst, 1):5.1f}")


if __name__ == "__main__":
    main()
