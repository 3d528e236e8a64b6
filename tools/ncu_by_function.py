"""Aggregate an ncu source-page export per device function.

  ncu -i X.ncu-rep --page source --csv > src.csv
  python tools/ncu_by_function.py src.csv monsoon_b200/libsb_b200.so [kernel-substring]

The SASS page has addresses only; the noinline device functions of a kernel are `$kernel$func` symbols of
the kernel's .text section (cuobjdump -elf), so address - first address = section offset -> function.
Columns: share of executed warp instructions, average active threads, share of stall samples, and the
split of the function's samples over the main stall reasons.
"""
import csv
import collections
import re
import subprocess
import sys


def symbols(so, kernel_mangled):
    out = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
    syms = []
    for line in out.splitlines():
        m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|0)\s+(0x[0-9a-f]+|0)\s+0x2\s+\S+\s+\S+\s+\$(\S+?)\$(\S+)\s*$", line)
        if m and m.group(3) == kernel_mangled:
            syms.append((int(m.group(1), 16), int(m.group(2), 16), m.group(4)))
    return sorted(set(syms))


def calibrate(body, syms):
    """Section offset of the first listed instruction.  The page does not always start at section offset 0: every
    CALL.REL target is a function start, so the shift is the one that maps most call targets onto symbol values."""
    a0 = int(body[0][0], 16)
    targets = set()
    for r in body:
        m = re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", r[1])
        if m:
            targets.add(int(m.group(1), 16) - a0)
    starts = {v for v, _sz, _n in syms}
    best, best_hits = 0, -1
    for t in targets:
        for v in starts:
            d = v - t
            hits = sum(1 for x in targets if x + d in starts)
            if hits > best_hits:
                best, best_hits = d, hits
    if targets and best_hits < 0.9 * len(targets):
        sys.stderr.write("WARNING: only %d of %d call targets land on function symbols -- is this the library the capture "
                         "was taken with?\n" % (best_hits, len(targets)))
    return best


def demangle(n):
    m = re.match(r"_Z(\d+)", n)
    return n[m.end():m.end() + int(m.group(1))] if m else n


def main():
    src, so = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    kname = rows[0][1]
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    # mangled name of the kernel: find the .text section whose demangled template args match
    out = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
    base = re.match(r"void (\w+)", kname).group(1) if kname.startswith("void") else kname.split("(")[0]
    targs = re.findall(r"\((?:bool|int)\)(\d+)", kname.split(">(")[0]) if "<" in kname else []
    cands = sorted(set(re.findall(r"\.text\.(_Z\d+%s\S*)" % base, out)))
    pick = None
    for c in cands:
        got = re.findall(r"L[bi](\d+)E", c)[:len(targs)] if targs else []
        if got == targs:
            pick = c
    if pick is None:
        pick = cands[0]
    syms = symbols(so, pick)
    body = rows[2:]
    a0 = int(body[0][0], 16) - calibrate(body, syms)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: collections.Counter())
    for r in body:
        off = int(r[0], 16) - a0
        fn = "(kernel body)"
        for v, sz, name in syms:
            if v <= off < v + sz:
                fn = demangle(name)
                break
        a = agg[fn]
        a["inst"] += int(r[col["Instructions Executed"]] or 0)
        a["thr"] += int(r[col["Thread Instructions Executed"]] or 0)
        a["samples"] += int(r[col["# Samples"]] or 0)
        a["l2local"] += int(r[col["L2 Theoretical Sectors Local"]] or 0)
        for s in stall_cols:
            a[s] += int(r[col[s]] or 0)
    ti = sum(a["inst"] for a in agg.values())
    ts = sum(a["samples"] for a in agg.values())
    tl = sum(a["l2local"] for a in agg.values()) or 1
    print("kernel", kname[:100])
    print("section", pick, "| warp instructions %.3e | samples %d" % (ti, ts))
    print("%-24s %7s %6s %8s %8s | %s" % ("function", "inst%", "thr", "sample%", "l2loc%", "top stall reasons (share of the function's samples)"))
    for fn, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        if a["samples"] * 200 < ts:
            continue
        top = sorted(((a[s], s) for s in stall_cols), reverse=True)[:4]
        print("%-24s %6.1f%% %6.2f %7.1f%% %7.1f%% | %s" % (
            fn[:24], 100.0 * a["inst"] / ti, a["thr"] / max(a["inst"], 1), 100.0 * a["samples"] / ts, 100.0 * a["l2local"] / tl,
            ", ".join("%s %.0f%%" % (s[6:], 100.0 * v / max(a["samples"], 1)) for v, s in top)))


if __name__ == "__main__":
    main()
