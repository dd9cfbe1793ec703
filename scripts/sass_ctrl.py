"""Decode the scheduling control bits of sm_100 SASS (stall count, write/read barrier, wait mask) from `cuobjdump -sass`.
Usage: python scripts/sass_ctrl.py <obj> <mangled-name-pattern> [start_hex end_hex]"""
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
names = subprocess.run("cuobjdump -elf %s | grep -o '_ZN2mb[A-Za-z0-9_]*' | grep -E '%s' | grep -v _param_ | sort -u | head -1" % (obj, pat),
                       shell=True, capture_output=True, text=True).stdout.split()
txt = subprocess.run(["cuobjdump", "-sass", "-fun", names[0], obj], capture_output=True, text=True).stdout.splitlines()
i = 0
while i < len(txt):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", txt[i])
    if m and i + 1 < len(txt):
        m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", txt[i + 1])
        addr = int(m.group(1), 16)
        if m2 and lo <= addr <= hi:
            w = int(m2.group(1), 16)  # upper 64 bits: bit k of the instruction = bit k-64 here
            stall = (w >> (105 - 64)) & 0xf
            yld = (w >> (109 - 64)) & 1
            wb = (w >> (110 - 64)) & 7
            rb = (w >> (113 - 64)) & 7
            wait = (w >> (116 - 64)) & 0x3f
            print("%05x  st%-2d %s wr:%s rd:%s wait:%s  %s" % (addr, stall, "Y" if yld else " ", "-" if wb == 7 else wb, "-" if rb == 7 else rb,
                                                           "".join(str(k) if wait >> k & 1 else "." for k in range(6)), m.group(2).strip()))
        i += 2
    else:
        i += 1
