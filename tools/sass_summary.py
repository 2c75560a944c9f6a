"""Static evidence from the shipped library (no GPU needed): per-kernel resource usage (cuobjdump -res-usage) and the counts
of the SASS mnemonics that identify the Blackwell paths (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld),
UTMALDG (TMA tensor loads), UBLKCP (cp.async.bulk), SYNCS (mbarrier), DMMA (FP64 tensor), UTCBAR (tcgen05.commit).
Usage: python tools/sass_summary.py > profiles/sass_<tag>.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyarrowspace_b200", "libarrowspace_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "UBLKCP", "UBLKPF", "SYNCS", "DMMA", "DFMA", "DMUL", "DADD", "HMMA", "REDUX",
             "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    name = name.replace("(anonymous namespace)::", "")
    name = re.sub(r"^void ", "", name)
    depth = 0
    for i, ch in enumerate(name):                          # drop the argument list: first '(' outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return name[:i]
    return name


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(v) for v in m.groups())
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    totals = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op = m.group(1)
            if op in MNEMONICS:
                counts[cur][op] += 1
                totals[op + ("" if op != "UTCHMMA" else (".2CTA" if ".2CTA" in m.group(2) else ""))] += 1
    names = demangle(sorted(set(usage) | set(counts)))
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout +
                                  sass)))
    print("# SASS summary of `pyarrowspace_b200/libarrowspace_b200.so` (static, `tools/sass_summary.py`)\n")
    print("Architectures in the fat binary: %s.  %d kernels.\n" % (", ".join(archs) or "?", len(usage)))
    print("Totals: " + ", ".join("%s x%d" % kv for kv in sorted(totals.items())) + "\n")
    print("| kernel | regs | stack B | static smem B | local B | Blackwell / FP64 mnemonics |")
    print("|---|---|---|---|---|---|")
    for fn in sorted(usage, key=lambda k: short(names[k])):
        reg, stack, shared, local = usage[fn]
        c = counts.get(fn, {})
        ops = ", ".join("%s x%d" % (k, c[k]) for k in MNEMONICS if c.get(k) and k not in ("DFMA", "DMUL", "DADD"))
        print("| `%s` | %d | %d | %d | %d | %s |" % (short(names[fn])[:110], reg, stack, shared, local, ops))
    spills = [short(names[f]) for f, u in usage.items() if u[1] > 0 or u[3] > 0]
    print("\nKernels with a stack frame or local memory: %s" % (", ".join("`%s`" % s for s in sorted(set(spills))) or "none"))


if __name__ == "__main__":
    sys.exit(main())
