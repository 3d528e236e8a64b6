"""dev: SASS of one out-of-line device function inside a kernel, by the `$kernel$function` symbols of the cubin.
  python tools/sass_func.py lib.so <kernel mangled name> [function substring | --table]"""
import os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_by_function import symbols, demangle
so, kern = sys.argv[1], sys.argv[2]
what = sys.argv[3] if len(sys.argv) > 3 else "--table"
syms = symbols(so, kern)
sass = subprocess.run(["cuobjdump", "-sass", "-fun", kern, so], capture_output=True, text=True).stdout
ins = []
for line in sass.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
def rng(lo, hi): return [(a, s) for a, s in ins if lo <= a < hi]
if what == "--table":
    first = min(v for v, _s, _n in syms) if syms else 1 << 30
    rows = [("(kernel body)", rng(0, first))] + [(demangle(n), rng(v, v + sz)) for v, sz, n in syms]
    print("%-26s %6s %5s %5s %5s %5s %5s" % ("function", "instr", "STL", "LDL", "LDS", "STS", "CALL"))
    for name, body in sorted(rows, key=lambda r: -len(r[1])):
        c = lambda pat: sum(1 for _a, s in body if re.search(pat, s))
        print("%-26s %6d %5d %5d %5d %5d %5d" % (name, len(body), c(r"\bSTL"), c(r"\bLDL"), c(r"\bLDS"), c(r"\bSTS"), c(r"\bCALL")))
else:
    for v, sz, n in syms:
        if what in n:
            for a, s in rng(v, v + sz): print("%6x  %s" % (a - v, s))
