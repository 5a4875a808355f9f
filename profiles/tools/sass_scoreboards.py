"""Decode the scoreboard fields of sm_100 SASS control words (cuobjdump -sass output of one function on
stdin or argv[1]): stall count, write / read barrier index, wait mask.  Used to find where ptxas
parks the wait for a long-latency load (profiles/README.md)."""
import re
import sys

lines = (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin).read().split('\n')
lo_addr = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi_addr = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
i = 0
while i < len(lines):
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/', lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r'\s+/\* 0x([0-9a-f]{16}) \*/', lines[i + 1])
        if m2:
            w = (int(m2.group(1), 16) << 64) | int(m.group(3), 16)
            a = int(m.group(1), 16)
            stall, wb, rb, wait = (w >> 105) & 0xf, (w >> 110) & 7, (w >> 113) & 7, (w >> 116) & 0x3f
            if lo_addr <= a <= hi_addr:
                print(f"{a:05x} st={stall:2d} wb={wb if wb != 7 else '-'} rb={rb if rb != 7 else '-'} wait={wait:06b}  {m.group(2).strip()[:90]}")
            i += 2
            continue
    i += 1
