"""Regenerate profiles/r02_plan_kernel_sass.md from the built library (cuobjdump -sass; no GPU needed).

    python scripts/sass_excerpt.py > profiles/r02_plan_kernel_sass.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rrt_star_fnd_algorithm_b200", "librrtfnd.so")
NAME = "_ZN6rrtfnd11plan_kernelILi1ELi16ELb0ELi1ELb1ELb0ELb0ELb0EEEv13rrtfnd_params12rrtfnd_batch"

txt = subprocess.run(["cuobjdump", "-sass", "-fun", NAME, LIB], capture_output=True, text=True).stdout
lines = []
for l in txt.split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?;)", l)
    if m:
        lines.append((int(m.group(1), 16), m.group(2).rstrip()))
if not lines:
    sys.exit("kernel not found in " + LIB)


def opcode(s):
    t = s.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].rstrip(";")


ops = collections.Counter(opcode(s).split(".")[0] for _, s in lines)
full = collections.Counter(opcode(s) for _, s in lines)


def count(prefix):
    return sum(v for k, v in full.items() if k.startswith(prefix))


def excerpt(lo, hi):
    return "\n".join(f"        /*{a:04x}*/    {s}" for a, s in lines if lo <= a <= hi)


def find(pred, start=0):
    for i, (a, s) in enumerate(lines):
        if i >= start and pred(s):
            return i
    return -1


print("# Round 2 - SASS of the shipped planner instance (`cuobjdump -sass`, excerpt; regenerate with `python scripts/sass_excerpt.py`)\n")
print("`plan_kernel<WPB=1, MINW=16, NG=false, MODE=RRT*, RING=true, WREC=false, RECT=false, XORD=false>`")
print(f"(`{NAME}`): what `python bench.py` (c3, 16,384 planners) launches eight times per step. Built by")
print("`__graft_entry__.build()` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a --fmad=false --prec-div=true --prec-sqrt=true -lineinfo`).\n")
print(f"Instructions: {len(lines)}; 128 registers, 0 barriers (ptxas). Opcode histogram (top 24): "
      + ", ".join(f"{k} {v}" for k, v in ops.most_common(24)) + ".\n")
print(f"Of note: `UBLKCP` {count('UBLKCP')} (1-D TMA bulk copy global -> shared), `SYNCS.ARRIVE.TRANS64` {count('SYNCS.ARRIVE')} / "
      f"`SYNCS.PHASECHK` {count('SYNCS.PHASECHK')} (mbarrier expect-tx / try-wait), `LDS.128` {count('LDS.128')}, `DADD` {count('DADD')} / "
      f"`DMUL` {count('DMUL')} / `DFMA` {count('DFMA')} (separately rounded adds and multiplies; the `DFMA`s are the one fused multiply-add of "
      f"`calc_distance` and the Newton steps inside IEEE division / square root), `DSETP` {count('DSETP')}, `REDUX` {count('REDUX')} (warp reductions of the argmin), "
      f"`MUFU.RCP64H` {count('MUFU.RCP64H')} / `MUFU.RSQ64H` {count('MUFU.RSQ64H')} (seeds of `__ddiv_rn` / `__dsqrt_rn`), `PREEXIT` {count('PREEXIT')} + "
      f"`NANOSLEEP` {count('NANOSLEEP')} (sliced launches: `griddepcontrol.launch_dependents`, ticket wait), `STL` {count('STL')} / `LDL` {count('LDL')} (spills). "
      "No tensor-core instruction: the path has no contraction.\n")

i = find(lambda s: "NANOSLEEP" in s)
print("## Sliced-launch prologue (ticket acquire spin / release, launch_dependents)\n\n```")
print(excerpt(lines[max(i - 12, 0)][0], lines[find(lambda s: "PREEXIT" in s)][0]))
print("```\n")

i = find(lambda s: "UBLKCP" in s)
j = find(lambda s: "SYNCS.PHASECHK" in s)
print("## Nearest scan: wait for the tile (mbarrier try-wait), LDS.128 per lane, refill the stage with one elected lane's TMA bulk copy\n\n```")
print(excerpt(lines[j][0], lines[i + 2][0]))
print("```\n")

# the circle test inside seg_blocked: delta, then the root filter, then (rarely) the square root and the quotients
k = find(lambda s: "7.1054273576010018587e-15" in s)
print("## Circle test in `seg_blocked`: `b`, `c`, `delta`; the root filter (2^-47 = 7.105e-15: `E = S * 2^-47`; squares of `|L| +- E`, `|H| +- E` against `delta`);\n"
      "## only when that is too close to call, the calls into `__dsqrt_rn` and the paired `__ddiv_rn`, and the reference's four comparisons\n\n```")
start = k
while start > 0 and "LDG.E.128" not in lines[start][1]:
    start -= 1
end = find(lambda s: "CALL.REL" in s, k)
end2 = find(lambda s: "CALL.REL" in s, end + 1)
print(excerpt(lines[start - 1][0], lines[min(end2 + 12, len(lines) - 1)][0]))
print("```")
