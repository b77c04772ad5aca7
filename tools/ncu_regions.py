"""Attribute an ncu source-page export of step_kernel<true> to REGIONS of the kernel (outermost inline frame), and report
executed warp-instructions, stall samples, mean active threads and static SASS bytes per region.
usage: python tools/ncu_regions.py <rep> <libh1v2_b200.so>"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, so = sys.argv[1], sys.argv[2]
KSEL = sys.argv[3] if len(sys.argv) > 3 else "step_kernelILb1ELb0"  # cubin section of the profiled instantiation (CaT: step_kernelILb1ELb1)
CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "h1v2_isaac_b200", "csrc")


def sections(fname, start_pat):
    """Regions = the '// ---- title' section comments of the source, from the function matching start_pat to its end."""
    out, on, lines = [], False, open(os.path.join(CSRC, fname)).read().split("\n")
    for i, l in enumerate(lines, 1):
        if re.search(start_pat, l):
            on = True
            out.append([i, None, "prologue"])
        m = re.match(r"\s*// ---- (.*?)( ----)?$", l)
        if on and m:
            out[-1][1] = i - 1
            out.append([i, None, m.group(1)[:44]])
    if out:
        out[-1][1] = len(lines)
    return out


PHYS = sections("h1v2_physics.cuh", r"void substep\(")
STEP = sections("h1v2_step.cuh", r"step_kernel\(const")
EMIT = sections("h1v2_step.cuh", r"void emit_observation\(")
PHYS0 = PHYS[0][0]


def region(chain):
    # chain: innermost -> outermost list of (file, line); attribute to the OUTERMOST frame inside substep(), else step_kernel()
    for f, ln in reversed(chain):
        if f == "h1v2_physics.cuh" and ln >= PHYS0:
            for a, b, n in PHYS:
                if a <= ln <= b:
                    return "phys: " + n
    for f, ln in reversed(chain):
        if f == "h1v2_step.cuh":
            for a, b, n in STEP:
                if a <= ln <= b:
                    return "step: " + n
            return "step: helpers (obs/reset/command)"
    return "other: " + (chain[-1][0] if chain else "?")


tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(so)} > /dev/null", shell=True, check=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "config" not in f][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
amap, chain, insec, fresh = {}, [], False, True
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        insec = KSEL in m.group(1)
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
STALLS = ["stall_no_inst", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_branch_resolving", "stall_selected", "stall_math", "stall_mio", "stall_lg", "stall_dispatch"]
isx = [hdr.index(c) for c in STALLS]
stall = collections.defaultdict(lambda: [0] * len(STALLS))
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
    for q, ix in enumerate(isx):
        stall[k][q] += int(r[ix] or 0)
    tot[0] += e; tot[1] += s; tot[2] += t
print(f"total warp-instr {tot[0]:,}  samples {tot[1]:,}  avg threads {tot[2]/max(tot[0],1):.1f}  static SASS {sum(static.values()):,} B")
print(f"{'region':34s} {'instr%':>7s} {'samples%':>9s} {'thr/inst':>9s} {'SASS bytes':>11s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    top = sorted(zip(stall[k], STALLS), reverse=True)[:3]
    ts = " ".join(f"{n[6:]}={100*c/max(v[1],1):.0f}%" for c, n in top)
    print(f"{k[:34]:34s} {100*v[0]/tot[0]:7.1f} {100*v[1]/tot[1]:9.1f} {v[2]/max(v[0],1):9.1f} {static[k]:11,d}  {ts}")
