"""Per-instruction execution counts of ONE device function from an ncu source-page export.

  python tools/ncu_function_sass.py src.csv lib.so <kernel .text section mangled name> <function substring>
Columns: offset, executions per call (first instruction = calls), average active threads, SASS.
"""
import csv, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_by_function import calibrate, symbols
src, so, section, fn = sys.argv[1:5]
rows = list(csv.reader(open(src)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
a0 = int(body[0][0], 16) - calibrate(body, symbols(so, section))
out = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
sym = None
for line in out.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)\s*$", line)
    if m and m.group(3) == section and fn in m.group(4):
        sym = (int(m.group(1), 16), int(m.group(2), 16), m.group(4))
print(sym)
lines = []
for r in body:
    off = int(r[0], 16) - a0
    if sym[0] <= off < sym[0] + sym[1]:
        lines.append((off - sym[0], int(r[col["Instructions Executed"]] or 0), int(r[col["Thread Instructions Executed"]] or 0), r[1].strip()))
calls = max(lines[0][1], 1)
print("static instructions", len(lines), "calls", calls, "warp instructions per call %.1f" % (sum(l[1] for l in lines) / calls))
for off, n, t, s in lines:
    print("%5x %7.2f %5.2f  %s" % (off, n / calls, t / max(n, 1), s[:100]))
