import csv, subprocess, sys, io, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; c = {n: i for i, n in enumerate(h)}
ops = collections.Counter(); samp = collections.Counter(); lines = []
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) < len(h): continue
    sass = r[c["Source"]].strip()
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    op = op.split(".")[0]
    n = int(float(r[c["Instructions Executed"]] or 0)); s = int(float(r[c["# Samples"]] or 0))
    ops[op] += n; samp[op] += s; tot_i += n; tot_s += s
    lines.append((s, n, sass))
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print("-- by opcode: %instr  %samples")
for op, n in ops.most_common(topn):
    print(f"{op:12s} {100*n/tot_i:5.1f}%  {100*samp[op]/max(tot_s,1):5.1f}%")
print("-- top SASS lines by samples")
for s, n, sass in sorted(lines, reverse=True)[:topn]:
    print(f"{100*s/max(tot_s,1):5.1f}%  exec {n:9d}  {sass[:100]}")
