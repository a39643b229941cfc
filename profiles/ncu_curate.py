#!/usr/bin/env python3
"""One entry per kernel family out of the raw CSV of an `ncu --set full` pass over profiles/prof_all.py:
summary text (as ncu_summary.py prints it), the per-kernel DRAM-throughput table and the traffic JSON bench.py reads.
usage: python profiles/ncu_curate.py raw.csv TAG      -> profiles/TAG_ncu_summary.txt, TAG_kernel_rooflines.md, TAG_traffic.json"""
import csv
import io
import json
import os
import sys
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ncu_summary  # noqa: E402

# (launch index in the pass, 1-based, label)
PICK = [(8, "cfg2, 3rd step"), (9, "cfg2, 3rd step"), (10, "cfg2, fpx_step_dev: + replica handleChosen + watermark, 3rd step"),
        (13, "cfg2, stand-alone fpx_proxyleader_phase2b_dev"), (14, ""), (15, ""), (16, ""), (17, ""), (18, ""),
        (20, "cfg5 vm_client_request: arm + own vote"), (21, "cfg5, 6*2^20 messages"), (22, "cfg5 vanilla Mencius: blind stamps + row sweep"),
        (25, "cfg4, 384-byte leader rows written by whole warps"), (26, "cfg4 PreAccept"), (28, ""),
        (30, "cfg4 PreAcceptOk, stamp pass"), (31, "cfg4 PreAcceptOk, decide pass"), (32, ""), (36, "conflict index, first pass"),
        (46, ""), (48, ""), (52, "")]


def main(path, tag):
    rows = list(csv.reader(open(path).read().splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    col = {h: i for i, h in enumerate(hdr)}
    peak = json.load(open(os.path.join(os.path.dirname(HERE), "MEASURED_PEAKS.json")))["hbm_gbs"]
    out_txt, table, traffic = [], [], {}
    for idx, label in PICK:
        r = body[idx - 1]
        buf = io.StringIO()
        tmp = io.StringIO()
        w = csv.writer(tmp)
        w.writerow(hdr); w.writerow(units); w.writerow(r)
        p = "/tmp/_one.csv"
        open(p, "w").write(tmp.getvalue())
        with redirect_stdout(buf):
            ncu_summary.main(p)
        text = buf.getvalue().replace("fpx::", "")
        if label:
            first, rest = text.split("\n", 1)
            text = f"{first}   [{label}]\n{rest}"
        out_txt.append(text)
        f = lambda name: float(r[col[name]].replace(",", ""))
        us, rd, wr = f("gpu__time_duration.sum"), f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
        if units[col["dram__bytes_read.sum"]].lower().startswith("gbyte"):
            rd *= 1e3
        if units[col["dram__bytes_write.sum"]].lower().startswith("gbyte"):
            wr *= 1e3
        if units[col["dram__bytes_read.sum"]].lower().startswith("kbyte"):
            rd /= 1e3
        if units[col["dram__bytes_write.sum"]].lower().startswith("kbyte"):
            wr /= 1e3
        l2 = f("lts__t_sectors.sum") * 32 / 1e6
        name = r[ki].replace("fpx::", "").replace("void ", "")[:70]
        gbs = (rd + wr) / us * 1e3          # MB / us = TB/s
        table.append(f"| `{name.split('(')[0]}` | {label} | {us:.1f} | {rd + wr:.1f} | {gbs:.0f} | {gbs / peak:.2f} | {l2:.0f} |")
        key = name.split("(")[0].split("<")[0]
        if idx in (8, 10):
            traffic[key] = {"dram_bytes_read": rd * 1e6, "dram_bytes_write": wr * 1e6, "l2_bytes": l2 * 1e6, "us_under_ncu": us,
                            "case": label}
    base = os.path.join(HERE, tag)
    open(base + "_ncu_summary.txt", "w").write(
        "ncu --set full --clock-control none --import-source on (profiles/prof_all.py, one B200; every launch replayed ~40x with cold\n"
        "caches: durations here are NOT bench numbers -- compare shares and counters).  One entry per kernel family, picked by\n"
        "profiles/ncu_curate.py from the raw CSV export of the report (the .ncu-rep of the whole pass exceeds gpurun's 64 MiB).\n"
        "lts__t_sectors x 32 B = L2 traffic.\n\n" + "".join(out_txt))
    open(base + "_kernel_rooflines.md", "w").write(
        "# Per-kernel DRAM throughput under ncu\n\nDRAM bytes moved / the launch's duration under `ncu --set full` (cold caches, one launch each), against the measured HBM\n"
        f"copy peak of `MEASURED_PEAKS.json` ({peak} GB/s).  TRAFFIC rates under the profiler, not the bench's algorithmic-bytes rooflines.\n\n"
        "| Kernel | case | µs (ncu) | DRAM MB | DRAM GB/s | of measured peak | L2 MB |\n|---|---|---:|---:|---:|---:|---:|\n" + "\n".join(table) + "\n")
    json.dump(traffic, open(base + "_traffic.json", "w"), indent=1)
    print("\n".join(table))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
