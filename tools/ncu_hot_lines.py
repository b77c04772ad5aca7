"""Aggregate an ncu source-page export by source line (needs the matching .so for nvdisasm line info)."""
import csv, re, subprocess, sys, collections, os, tempfile
rep, so = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
KSEL = sys.argv[4] if len(sys.argv) > 4 else "step_kernelILb1ELb0"
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(so)} > /dev/null", shell=True, check=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "config" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
amap, cur, insec = {}, None, False
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        insec = KSEL in m.group(1)
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), os.path.basename(m.group(3)) if m.group(3) else None, int(m.group(4)) if m.group(4) else None)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", l)
    if m and cur:
        amap[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ie, isamp, ithr = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    try:
        a = int(r[ia], 16)
    except Exception:
        continue
    if base is None:
        base = a
    key = amap.get(a - base)
    if key is None:
        k2 = ("?", 0)
    else:
        f, ln, inf, inl = key
        k2 = (inf, inl) if inf in ("h1v2_physics.cuh", "h1v2_step.cuh") and f not in ("h1v2_physics.cuh", "h1v2_step.cuh") else (f, ln)
    e, s, t = int(r[ie] or 0), int(r[isamp] or 0), int(r[ithr] or 0)
    agg[k2][0] += e; agg[k2][1] += s; agg[k2][2] += t
    tot[0] += e; tot[1] += s; tot[2] += t
print(f"total warp-instr {tot[0]:,}  samples {tot[1]:,}  avg threads {tot[2]/max(tot[0],1):.1f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]:22s} {k[1]:5d}  instr {100*v[0]/tot[0]:5.1f}%  samples {100*v[1]/tot[1]:5.1f}%  thr/inst {v[2]/max(v[0],1):5.1f}")
