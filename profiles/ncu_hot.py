#!/usr/bin/env python3
"""Hot SASS instructions of one kernel in an .ncu-rep (stall samples + executed count).
usage: python profiles/ncu_hot.py rep.ncu-rep <kernel-regex> [top]"""
import csv
import subprocess
import sys


def main(path, kern, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # first launch only
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    start = hdr_i[0]
    end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
    hdr = rows[start]
    body = rows[start + 1:end]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    tot_s = sum(int(r[si]) for r in body if len(r) > ii)
    tot_i = sum(int(r[ii]) for r in body if len(r) > ii)
    print(f"{rows[start-1][1][:60]}  instructions={len(body)}  warp-inst executed={tot_i}  samples={tot_s}")
    ranked = sorted(((int(r[si]), k) for k, r in enumerate(body) if len(r) > ii), reverse=True)[:top]
    for s, k in sorted(ranked, key=lambda x: x[1]):
        r = body[k]
        print(f"  #{k:4d} {100.0*s/max(tot_s,1):5.1f}%  exec={int(r[ii]):8d}  {r[1].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
