"""Instruction mix and hottest SASS of one kernel from an .ncu-rep source page.
Usage: python benchmarks/ncu_source.py file.ncu-rep kernel_regex [topN]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    body = []
    for r in rows:
        if r and r[0] == "Address":
            if hdr is not None:
                break  # first launch only
            hdr = r
            continue
        if hdr is not None and len(r) == len(hdr):
            body.append(r)
    ci, si, ss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[ci]) for r in body)
    tots = sum(int(r[ss]) for r in body)
    print("total warp-inst", tot, "sass lines", len(body), "stall samples", tots)
    c, st = Counter(), Counter()
    for r in body:
        toks = r[si].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        c[op] += int(r[ci])
        st[op] += int(r[ss])
    for op, n in c.most_common(top):
        print("%-10s %9d %5.1f%%  stall %5.1f%%" % (op, n, 100 * n / tot, 100 * st[op] / max(tots, 1)))
    print("-- hottest stall lines")
    for r in sorted(body, key=lambda r: -int(r[ss]))[:top]:
        print("%6s %8s  %s" % (r[ss], r[ci], r[si].strip()[:110]))


if __name__ == "__main__":
    main()
