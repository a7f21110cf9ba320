"""Static instruction mix of the fp32 screen's hot loop (the 8-candidate batch loop) of a built library; no GPU needed.
usage: python tools/sass_screen_loop.py <lib.so> <kernel-substring> [min_ex2] [-v]
Finds backward branches and takes the shortest loop body with at least min_ex2 (default 24: the 8-candidate loop of the 3-points-per-thread shape) MUFU.EX2 (one per screened point evaluation)."""
import collections
import re
import subprocess
import sys

so, pat = sys.argv[1], sys.argv[2]
MIN_EX = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 24
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
            nex = sum(1 for _, s in seg if "MUFU.EX2" in s)
            if nex >= MIN_EX and (best is None or len(seg) < len(best[1])):  # the innermost loop with a whole candidate batch in it
                best = (nex, seg)
nex, seg = best
ops = collections.Counter()
for _, s in seg:
    t = s.split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += 1
print(f"{body[0].split(chr(10), 1)[0][:90]}")
print(f"  screen loop: {len(seg)} instructions for {nex} point evaluations = {len(seg) / nex:.2f} per evaluation")
print("  per evaluation: " + "  ".join(f"{k} {v / nex:.2f}" for k, v in ops.most_common()))
if "-v" in sys.argv:
    for a, s in seg:
        print(f"   {a:05x}  {s}")
