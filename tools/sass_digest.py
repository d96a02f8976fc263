"""Developer tool: per-kernel digest of the shipped library's SASS (registers, shared memory, instruction counts of
the mnemonics that prove the Blackwell-native paths) -> profiles/sass_digest.txt.
    python tools/sass_digest.py [gsdf_slam_b200/libgsr.so] [profiles/sass_digest.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gsdf_slam_b200", "libgsr.so")
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "sass_digest.txt")
WATCH = ["UBLKCP", "SYNCS", "LDGSTS", "RED", "REDG", "ATOMG", "ATOMS", "MUFU", "FFMA2", "REDUX", "SHFL", "VOTE", "MATCH", "LDS", "STS",
         "LDG", "STG", "BAR", "UTMALDG", "HMMA", "UTCHMMA"]

res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        usage[cur] = tuple(int(x) for x in m.groups())
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.defaultdict(collections.Counter)
total = collections.Counter()
archs = set(re.findall(r"arch = (sm_\w+)", sass))
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    out = {}
    for n, x in zip(names, r):
        x = re.sub(r"^void ", "", x).replace("gsr::", "").replace("(anonymous namespace)::", "")
        head = x.split("(")[0]
        out[n] = head if head else x
    return out


names = sorted(total)
dm = demangle(names)
lines = [f"# SASS digest of {os.path.relpath(lib, ROOT)} (cuobjdump -sass / -res-usage; architectures: {', '.join(sorted(archs))})",
         "# kernel | regs | stack B | static smem B | SASS instructions | watched mnemonics", ""]
for n in sorted(names, key=lambda x: dm[x]):
    reg, stack, smem, _ = usage.get(n, (0, 0, 0, 0))
    w = " ".join(f"{k}={v}" for k, v in sorted(counts[n].items()))
    lines.append(f"{dm[n]:58s} regs={reg:3d} stack={stack:3d} smem={smem:6d} inst={total[n]:5d}  {w}")
lines.append("")
lines.append("UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier operations, LDGSTS = cp.async, RED = fire-and-forget "
             "global reductions, REDUX = warp reductions into uniform registers; no HMMA/UTCHMMA: nothing on this path is a "
             "dense contraction.")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:12]))
