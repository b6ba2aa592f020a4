"""tools/sass_table.py [lib.so] -- per-kernel SASS census of libtess_b200.so (cuobjdump): instructions,
code size, registers, and the mnemonics that matter for this library (fp64 arithmetic, memory,
reductions, shuffles / votes, TMA: UTMAPF = tensor prefetch, UBLKCP = bulk copy; LDGSTS = cp.async).
Tensor cores are deliberately unused (no UTC*MMA / HMMA): a 2-D distance is not a contraction."""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tess_b200", "libtess_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", ln)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
GROUPS = [("fp64", ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX")), ("LDG", ("LDG",)), ("STG", ("STG",)), ("LDS", ("LDS",)),
          ("STS", ("STS",)), ("LDGSTS", ("LDGSTS",)), ("RED/ATOM", ("RED", "ATOM")), ("SHFL", ("SHFL",)),
          ("VOTE", ("VOTE",)), ("UTMAPF", ("UTMAPF",)), ("UBLKCP", ("UBLKCP",)), ("MUFU", ("MUFU",)), ("MMA", ("HMMA", "UTC", "DMMA", "IMMA"))]
print("| kernel | instr | KB | regs | smem B | " + " | ".join(g for g, _ in GROUPS) + " |")
print("|---|---|---|---|---|" + "---|" * len(GROUPS))
for f in re.split(r"\n\s*Function : ", sass)[1:]:
    name = f.split("\n")[0].strip()
    ops = []
    for l in f.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops.append(m.group(1))
    c = collections.Counter()
    for o in ops:
        for g, pre in GROUPS:
            if any(o.startswith(p) for p in pre):
                c[g] += 1
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
    r = regs.get(name, ("?", "?"))
    print("| `%s` | %d | %.1f | %s | %s | " % (short, len(ops), len(ops) * 16 / 1024.0, r[0], r[1]) + " | ".join(str(c[g]) for g, _ in GROUPS) + " |")
