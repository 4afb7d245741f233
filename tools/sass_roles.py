"""Spill / instruction-mix histogram of a kernel's SASS by 1600-line window (role attribution by eye: UBLKCP+UTCIMMA = service,
PRMT-heavy = slicing, LDTM = read-back, ...).  usage: sass_roles.py file.o mangled-kernel-name-substring"""
import subprocess, sys
from collections import Counter
o, name = sys.argv[1], sys.argv[2]
fn = [l.split()[-1].rstrip(":") for l in subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout.splitlines() if "Function :" in l and name in l][0]
lines = subprocess.run(["cuobjdump", "-sass", "-fun", fn, o], capture_output=True, text=True).stdout.splitlines()
keys = ["USETMAXREG", "UBLKCP", "UTCIMMA", "LDTM", "DMMA", "F2I.S64", "STL", "LDL"]
for s in range(0, len(lines), 1600):
    c = Counter(k for l in lines[s:s + 1600] for k in keys if k in l)
    print(s, dict(c))
