"""Static instruction mix of the hottest inner loop of a kernel (no GPU needed).
usage: python tools/sass_loop.py <lib.so> <kernel-substring> [min_fp64]
Finds backward branches, takes the shortest loop body with at least 20 fp64-pipe instructions."""
import collections
import re
import subprocess
import sys

so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
body = [b for b in blocks if pat in b.split("\n", 1)[0]]
if not body:
    sys.exit("kernel not found")
lines = []
for ln in body[0].split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(lines)}
best = None
for i, (a, ins) in enumerate(lines):
    m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", ins)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_index:
            seg = lines[addr_index[tgt]:i + 1]
            n64 = sum(1 for _, s in seg if re.match(r"(@!?U?P\d+\s+)?(DFMA|DMUL|DADD|DSETP)", s))
            nfma = sum(1 for _, s2 in seg if "DFMA" in s2)
            if nfma >= 15 and (best is None or len(seg) < len(best[1])):
                best = (n64, seg)
n64, seg = best
ops = collections.Counter()
for _, s in seg:
    t = s.split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += 1
print(f"{pat}: loop of {len(seg)} instructions, fp64-pipe {n64}")
print("  " + "  ".join(f"{k} {v}" for k, v in ops.most_common()))
if "-v" in sys.argv:
    for a, s in seg:
        print(f"   {a:05x}  {s}")
