"""SASS evidence for the Blackwell-specific instructions each kernel of libvfi_b200.so uses (cuobjdump -sass):
tcgen05 (UTCHMMA / UTCBAR / STTM / LDTM), bulk copies (UBLKCP), cp.async (LDGSTS), mbarriers (SYNCS), vector reductions
(REDG ... F32x4), HFMA2.BF16.  usage: sass_mnemonics.py [path/to/libvfi_b200.so]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "video-frame-interpolation_b200/libvfi_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = {"UTCHMMA": r"\bUTCHMMA", "UTCBAR": r"\bUTCBAR", "STTM": r"\bSTTM", "LDTM": r"\bLDTM", "UBLKCP": r"\bUBLKCP",
        "LDGSTS": r"\bLDGSTS", "SYNCS": r"\bSYNCS", "REDG.F32x4": r"\bREDG\.E\.ADD\.F32x4", "HFMA2.BF16": r"HFMA2\.BF16",
        "LDS.128": r"\bLDS\.128", "FENCE.VIEW.ASYNC": r"FENCE\.VIEW\.ASYNC", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG",
        "CREDUX": r"\bCREDUX", "LDS.U16": r"\bLDS\.U16"}
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        cur = re.sub(r"\(.*", "", name)[:110]
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
print(f"# {lib}: static SASS instruction counts per kernel (sm_100a)")
for name, c in counts.items():
    if any(c[k] for k in ("UTCHMMA", "STTM", "LDTM", "UBLKCP", "LDGSTS", "REDG.F32x4", "UTMALDG", "UTMASTG")):
        print(name)
        print("    " + "  ".join(f"{k}={v}" for k, v in c.items() if v))
