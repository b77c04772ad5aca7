"""Attribute an ncu source-page export of step_kernel<true> to REGIONS of the kernel (outermost inline frame), and report
executed warp-instructions, stall samples, mean active threads and static SASS bytes per region.
usage: python tools/ncu_regions.py <rep> <libh1v2_b200.so>"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, so = sys.argv[1], sys.argv[2]
PHYS = [(274, 310, "stage+root frame"), (311, 329, "pass1 sincos/ankle"), (330, 358, "pass2 kinematics+RNE"), (359, 389, "smooth forces/rows"),
        (390, 430, "contact candidates"), (431, 447, "iterate init"), (448, 514, "newton: evaluate rows"), (515, 521, "phase2 rhs"),
        (522, 555, "ABA sweep1"), (556, 588, "root 6x6"), (589, 614, "ABA sweep2"), (615, 620, "step-tol exit"), (621, 640, "M-product"),
        (641, 692, "line search"), (693, 710, "iterate update"), (711, 740, "integrate")]
STEP = [(134, 171, "obs: sample+noise"), (172, 212, "obs: flatten/emit"), (214, 256, "load state"), (257, 271, "action"),
        (272, 306, "PD+sensor loop"), (307, 331, "guards/terminations"), (332, 403, "rewards"), (404, 460, "epsum/diag/stats"),
        (461, 489, "reset"), (490, 509, "command/push"), (510, 533, "store state")]


def region(chain):
    # chain: innermost -> outermost list of (file, line)
    for f, ln in reversed(chain):
        if f == "h1v2_physics.cuh" and ln >= 274:
            for a, b, n in PHYS:
                if a <= ln <= b:
                    return "phys: " + n
        if f == "h1v2_step.cuh":
            hit = None
            for a, b, n in STEP:
                if a <= ln <= b:
                    hit = "step: " + n
            if hit and not (ln == 288):
                return hit
    for f, ln in reversed(chain):
        if f == "h1v2_step.cuh":
            for a, b, n in STEP:
                if a <= ln <= b:
                    return "step: " + n
    return "other: " + (chain[-1][0] if chain else "?")


tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(so)} > /dev/null", shell=True, check=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "config" not in f][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
amap, chain, insec, fresh = {}, [], False, True
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        insec = "step_kernelILb1" in m.group(1)
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", l)
    if m:
        amap[int(m.group(1), 16)] = list(chain)
        fresh = True
static = collections.Counter()
for a, ch in amap.items():
    static[region(ch)] += 16
if rep == "-":
    print(f"static SASS {sum(static.values()):,} B")
    for k, v in sorted(static.items(), key=lambda kv: -kv[1]):
        print(f"{k:34s} {v:11,d}")
    sys.exit(0)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ie, isamp, ithr = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
base, agg, tot = None, collections.defaultdict(lambda: [0, 0, 0]), [0, 0, 0]
for r in rows[2:]:
    try:
        a = int(r[ia], 16)
    except Exception:
        continue
    if base is None:
        base = a
    k = region(amap.get(a - base, []))
    e, s, t = int(r[ie] or 0), int(r[isamp] or 0), int(r[ithr] or 0)
    agg[k][0] += e; agg[k][1] += s; agg[k][2] += t
    tot[0] += e; tot[1] += s; tot[2] += t
print(f"total warp-instr {tot[0]:,}  samples {tot[1]:,}  avg threads {tot[2]/max(tot[0],1):.1f}  static SASS {sum(static.values()):,} B")
print(f"{'region':34s} {'instr%':>7s} {'samples%':>9s} {'thr/inst':>9s} {'SASS bytes':>11s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} {100*v[0]/tot[0]:7.1f} {100*v[1]/tot[1]:9.1f} {v[2]/max(v[0],1):9.1f} {static[k]:11,d}")
