"""Per-source-line shared-memory wavefronts of one kernel in an .ncu-rep (ncu --set full
--import-source on), normalised per output pixel.  usage: ncu_wavefronts.py REPORT KERNEL_REGEX PIXELS_PER_LAUNCH"""
import collections
import csv
import subprocess
import sys


def main():
    rep, regex, pixels = sys.argv[1], sys.argv[2], float(sys.argv[3])
    out = subprocess.run(["ncu", "-i", rep, "-k", f"regex:{regex}", "-c", "1", "--page", "source", "--csv", "--print-source",
                          "sass,cuda"], capture_output=True, text=True).stdout
    cur, ci = None, None
    wf, ideal, tags = collections.Counter(), collections.Counter(), collections.Counter()
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            ci = {n: i for i, n in enumerate(r)}
        elif ci is not None and r[0] and r[0] != "Function Name":
            def g(n):
                try:
                    return int(r[ci[n]] or 0)
                except (ValueError, KeyError, IndexError):
                    return 0
            key = f"{cur}:{r[0]}  {r[1].strip()[:100]}"
            wf[key] += g("L1 Wavefronts Shared")
            ideal[key] += g("L1 Wavefronts Shared Ideal")
            tags[key] += g("L1 Tag Requests Global")
    tot = sum(wf.values())
    print(f"# {rep}: kernel /{regex}/, {pixels:.0f} pixels per launch")
    print(f"# shared-memory wavefronts per pixel: {tot / pixels:.2f} (ideal for the access widths used: {sum(ideal.values()) / pixels:.2f})")
    print("# wavefronts/px  ideal/px  source line")
    for k, v in wf.most_common(16):
        if v:
            print(f"{v / pixels:10.2f} {ideal[k] / pixels:9.2f}  {k}")


if __name__ == "__main__":
    main()
