"""Aggregate ncu per-SASS-instruction samples by CUDA source line.
usage: python scripts/ncu_lines.py <report.ncu-rep> <kernel mangled-name substring> [top]
Needs nvdisasm/cuobjdump and the in-tree librrtfnd.so built from the same sources as the profiled run."""
import csv, os, re, subprocess, sys, tempfile, collections

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("RRTFND_SO", os.path.join(root, "rrt_star_fnd_algorithm_b200", "librrtfnd.so"))
by_inst = os.environ.get("BY_INST") is not None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
sass = None
for f in sorted(os.listdir(tmp)):
    out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kname in out:
        sass = out
        break
assert sass, "kernel not found"
# map instruction offset -> (file, line) inside the wanted function
addr2line, cur, infunc = {}, None, False
for ln in sass.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        infunc = kname in m.group(1)
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur:
        addr2line[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, isamp, iexec = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r in rows[2:]:
    if len(r) <= isamp or not r[ia]:
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    s = int(r[isamp] or 0)
    per[key][0] += s
    per[key][1] += int(r[iexec] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            per[key][2][hdr[i]] += v
    tot += s
src = {}
tot_inst = sum(v[1] for v in per.values())
print(f"total samples {tot}, total warp instructions {tot_inst}")
for (f, l), (s, ex, st) in sorted(per.items(), key=lambda kv: -(kv[1][1] if by_inst else kv[1][0]))[:top]:
    if f not in src:
        p = os.path.join(root, "rrt_star_fnd_algorithm_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    main = ",".join(f"{k[6:]}:{v}" for k, v in st.most_common(3))
    print(f"{100.0 * s / max(tot, 1):5.1f}% samp {100.0 * ex / max(tot_inst, 1):5.1f}% inst  {f}:{l:<4d} [{main}]  {text}")
