"""SASS audit of libh1v2_b200.so (no GPU needed): size and opcode histogram per kernel, the Blackwell-specific opcodes called out.
usage: python tools/sass_hist.py > profiles/r2_sass_hist.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "h1v2_isaac_b200", "libh1v2_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist, size = None, {}, {}
for l in txt.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter(); size[kern] = 0
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and kern:
        op = m.group(2)
        hist[kern][op.split(".")[0]] += 1
        hist[kern]["full:" + op] += 1 if op.split(".")[0] in ("UBLKCP", "SYNCS", "LDGSTS", "REDUX", "RED", "ATOMG", "SHFL", "MUFU") else 0
        size[kern] = int(m.group(1), 16) + 16
print(f"# SASS audit of {os.path.relpath(so, ROOT)} (cuobjdump -sass, sm_100a); instruction size 16 B")
for k in sorted(size, key=lambda k: -size[k]):
    h = hist[k]
    base = {o: c for o, c in h.items() if not o.startswith("full:")}
    n = sum(base.values())
    print(f"\n## {k}: {size[k]:,} B of SASS, {n:,} instructions")
    print("   top opcodes: " + ", ".join(f"{o} {c}" for o, c in sorted(base.items(), key=lambda kv: -kv[1])[:22]))
    special = {o[5:]: c for o, c in h.items() if o.startswith("full:") and c}
    if special:
        print("   data movement / sync / special: " + ", ".join(f"{o} {c}" for o, c in sorted(special.items())))
    fp = sum(base.get(o, 0) for o in ("FFMA", "FMUL", "FADD", "FMNMX", "FSEL", "FSETP", "FCHK", "MUFU"))
    print(f"   FP32 pipe (FFMA FMUL FADD FMNMX FSEL FSETP MUFU): {fp} = {100 * fp / max(n, 1):.0f} % of the static instructions; LDS/STS {base.get('LDS', 0)}/{base.get('STS', 0)}; "
          f"LDL/STL {base.get('LDL', 0)}/{base.get('STL', 0)}; UBLKCP (cp.async.bulk) {base.get('UBLKCP', 0)}; LDGSTS (cp.async) {base.get('LDGSTS', 0)}; HMMA/UTC*MMA {base.get('HMMA', 0)}/{sum(c for o, c in base.items() if o.startswith('UTC'))}")
