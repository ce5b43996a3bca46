"""Per-kernel SASS evidence for the Blackwell-native claim (B200_PROFILING.md, "What proves a Blackwell-native kernel"):
counts of the tcgen05 / TMEM / TMA mnemonics in every kernel of libctcasr.so, plus the first lines around each
kernel's first UTC*MMA.  `python tools/sass_summary.py > profiles/r2_sass_summary.txt` (no GPU needed)."""
import collections
import re
import subprocess
import sys

LIB = "ctc_asr_b200/libctcasr.so"
PAT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCCP", "SYNCS", "HMMA", "FFMA", "LDGSTS", "SHFL", "ST.E", "RED"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kernels[cur] = []
    elif cur is not None and "/*" in line:
        kernels[cur].append(line)
print("SASS mnemonic counts per kernel of %s (cuobjdump -sass; nvcc 12.9, sm_100a)" % LIB)
print("%-62s %s" % ("kernel", " ".join("%7s" % p for p in PAT)))
for k, lines in kernels.items():
    body = "\n".join(lines)
    print("%-62s %s" % (k[-62:], " ".join("%7d" % len(re.findall(r"\b%s" % re.escape(p), body)) for p in PAT)))
print()
for k, lines in kernels.items():
    idx = [i for i, l in enumerate(lines) if "UTCHMMA" in l]
    if not idx:
        continue
    print("---- %s: first tcgen05.mma and its neighbours" % k)
    for l in lines[max(0, idx[0] - 3):idx[0] + 4]:
        print("   ", re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.strip())[:150])
