"""One fwd+bwd step cut out of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per-launch durations, shares, and totals per kernel.
usage: python scripts/launch_summary.py <launches.csv> [step-index-from-the-end (default 1)]"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
back = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
c = {n: i for i, n in enumerate(rows[hi])}
body = [r for r in rows[hi + 1:] if len(r) > c["Metric Value"]]
L = [(r[c["Kernel Name"]], r[c["Grid Size"]], float(r[c["Metric Value"]].replace(",", "")) / 1e3) for r in body]
starts = [i for i, (n, _, _) in enumerate(L) if "transpose_kernel" in n]
if len(starts) < back + 1:
    sys.exit("not enough complete steps in the list")
seg = L[starts[-back - 1]:starts[-back]]
short = lambda n: n.replace("void ", "").replace("ckb::", "").replace("<unnamed>::", "").split("(")[0]
ours = [x for x in seg if not x[0].startswith("void at::") and "at::native" not in x[0]]
tot = sum(t for _, _, t in seg)
print(f"{len(seg)} launches ({len(ours)} ours), {tot:.1f} us total, {sum(t for _, _, t in ours):.1f} us ours\n")
print(f"{'#':>3s} {'kernel':48s} {'grid':>14s} {'us':>9s} {'share':>6s}")
for i, (n, g, t) in enumerate(seg):
    print(f"{i:3d} {short(n)[:48]:48s} {g:>14s} {t:9.1f} {100 * t / tot:5.1f}%")
agg = collections.OrderedDict()
for n, _, t in seg:
    k = short(n)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += t
print("\nby kernel:")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"    {k[:56]:56s} x{n:<3d} {t:9.1f} us {100 * t / tot:5.1f}%")
