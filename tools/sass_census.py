"""Opcode census of libsfm_b200.so per kernel (cuobjdump -sass): the Blackwell-native instructions
(UTCIMMA = tcgen05.mma, UTMALDG / UBLKCP = TMA, LDTM = tcgen05.ld, SYNCS = mbarrier, USETMAXREG) next
to the integer / fp64 work.  Usage: python tools/sass_census.py > profiles/<tag>_sass_census.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "sfm_opencv_b200", "libsfm_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist, arch = None, collections.OrderedDict(), set()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ["UTCIMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "SYNCS", "USETMAXREG", "VIMNMX3", "VIMNMX", "VIADDMNMX",
       "IMAD", "VOTE", "LDS", "DFMA", "DMUL", "POPC", "LOP3", "LDG", "STG"]
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} (arch: {', '.join(sorted(arch))}); static instruction counts")
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{k:>9s}" for k in KEY))
tot = collections.Counter()
for k, h in hist.items():
    print(f"{k[:58]:58s} {sum(h.values()):6d} " + " ".join(f"{h.get(x, 0):9d}" for x in KEY))
    tot.update(h)
print(f"{'ALL':58s} {sum(tot.values()):6d} " + " ".join(f"{tot.get(x, 0):9d}" for x in KEY))
