"""How much of a capture is call overhead: per device function, calls (RET count), executed warp instructions and
the part that is register save/restore around calls (STL/LDL relative to the stack pointer R1, BMOV).

  python tools/ncu_abi_overhead.py src.csv lib.so <kernel .text section mangled name>
"""
import collections, csv, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_by_function import calibrate, demangle, symbols
src, so, section = sys.argv[1:4]
rows = list(csv.reader(open(src)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
syms = symbols(so, section)
a0 = int(body[0][0], 16) - calibrate(body, syms)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for r in body:
    off = int(r[0], 16) - a0
    n = int(r[col["Instructions Executed"]] or 0)
    s = r[1].strip()
    tot += n
    fn = "(kernel body)"
    for v, sz, name in syms:
        if v <= off < v + sz:
            fn = demangle(name)
            break
    a = agg[fn]
    a[0] += n
    if re.search(r"(STL|LDL)[.\w]* (R\d+, )?\[R1(\+0x[0-9a-f]+)?\]", s) or s.startswith("BMOV"):
        a[1] += n
    if "RET.REL" in s:
        a[2] += n
print("total warp instructions %.3e; stack save/restore + BMOV %.1f %%" % (tot, 100.0 * sum(a[1] for a in agg.values()) / tot))
print("%-24s %7s %7s %9s %s" % ("function", "inst%", "abi%", "calls", "abi instr/call"))
for fn, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%-24s %6.1f%% %6.1f%% %9d %6.1f" % (fn[:24], 100.0 * a[0] / tot, 100.0 * a[1] / tot, a[2], a[1] / max(a[2], 1)))
