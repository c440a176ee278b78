"""Top SASS instructions by warp-stall samples from an .ncu-rep (source page), with neighbours."""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
c = {n: i for i, n in enumerate(h)}
body = rows[hi + 1:]
tot = sum(int(r[c["# Samples"]] or 0) for r in body)
stalls = [n for n in h if n.startswith("stall_") and "(Not Issued)" not in n]
print(f"total samples {tot}")
order = sorted(range(len(body)), key=lambda i: -int(body[i][c["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    n = int(r[c["# Samples"]] or 0)
    st = sorted(((int(r[c[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{100*n/tot:5.1f}% [{i:5d}] {r[c['Source']].strip()[:70]:70s} x{r[c['Instructions Executed']]:>9s}  " +
          " ".join(f"{s}={v}" for v, s in st if v))
