#!/usr/bin/env python3
"""Attribute executed warp-instructions and stall samples of one kernel to CUDA
source lines: joins the SASS page of an .ncu-rep with `nvdisasm -g` line info of
the cubin inside libfpx.so (same build!).
usage: python profiles/ncu_lines.py rep.ncu-rep <kernel-regex> <mangled-substring> [top]"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(mangled_sub):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "frankenpaxos_b200/lib/libfpx.so")], cwd=d,
                   capture_output=True)
    cub = glob.glob(os.path.join(d, "*.cubin"))[0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.split("\n")
    out, cur, infn = [], None, False
    for ln in txt:
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            infn = mangled_sub in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            out.append(cur)
    return out


def main(rep, kern, mangled, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    start = hdr_i[0]
    end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
    hdr = rows[start]
    body = [r for r in rows[start + 1:end] if len(r) > 5]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    lines = sass_lines(mangled)
    if len(lines) != len(body):
        print(f"warning: {len(lines)} SASS lines in cubin vs {len(body)} in report (different build?)")
    agg_i, agg_s = collections.Counter(), collections.Counter()
    for k, r in enumerate(body):
        key = lines[k] if k < len(lines) else None
        agg_i[key] += int(r[ii]); agg_s[key] += int(r[si])
    ti, ts = sum(agg_i.values()), sum(agg_s.values())
    print(f"{kern}: warp-inst {ti}, samples {ts}")
    src = {}
    for key, v in sorted(agg_i.items(), key=lambda kv: -kv[1])[:top]:
        text = ""
        if key:
            f = key[0]
            if f not in src:
                p = os.path.join(ROOT, "frankenpaxos_b200/csrc", f)
                src[f] = open(p).read().split("\n") if os.path.exists(p) else []
            if src[f] and key[1] <= len(src[f]):
                text = src[f][key[1] - 1].strip()[:80]
        print(f"  {100*v/ti:5.1f}% inst {100*agg_s[key]/max(ts,1):5.1f}% stall  {str(key):34s} {text}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
