#!/usr/bin/env python
"""Warp-stall samples and executed instructions per source line of one kernel in an .ncu-rep (ncu --set full
--import-source on), with the dominant stall reasons.  usage: ncu_stalls.py REPORT KERNEL_REGEX [TOP]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, regex = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "-k", f"regex:{regex}", "-c", "1", "--page", "source", "--csv", "--print-source",
                          "sass,cuda"], capture_output=True, text=True).stdout
    cur, ci = None, None
    samples, inst, reasons = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            ci = {}
            for i, n in enumerate(r):
                ci.setdefault(n, i)
        elif ci is not None and r[0] and r[0] != "Function Name" and len(r) > ci.get("Address", 2) and r[ci["Address"]] == "-":
            def g(n):
                try:
                    return int(r[ci[n]] or 0)
                except (ValueError, KeyError, IndexError):
                    return 0
            key = f"{cur}:{r[0]}  {r[1].strip()[:90]}"
            samples[key] += g("# Samples")
            inst[key] += g("Instructions Executed")
            for n in ci:
                if n.startswith("stall_") and "Not Issued" not in n:
                    reasons[key][n[6:]] += g(n)
    tot, toti = sum(samples.values()), sum(inst.values())
    print(f"# {rep}: kernel /{regex}/: {tot} warp samples, {toti} warp instructions")
    print("# samples%  inst%   top stall reasons   source line")
    for k, v in samples.most_common(top):
        rs = ", ".join(f"{n} {c * 100 // max(v, 1)}%" for n, c in reasons[k].most_common(3))
        print(f"{100.0 * v / tot:7.2f} {100.0 * inst[k] / toti:6.2f}   [{rs}]   {k}")


if __name__ == "__main__":
    main()
