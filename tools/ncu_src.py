"""Top SASS instructions by warp-stall samples from an .ncu-rep (dev tool).
usage: python tools/ncu_src.py <report.ncu-rep> [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next((i for i, r in enumerate(rows) if "Source" in r and len(r) > 5), None)
if hi is None:
    print(out[:3000]); sys.exit(1)
hdr = rows[hi]; rows = rows[hi + 1:]
ci = {c: i for i, c in enumerate(hdr)}
cands = [c for c in hdr if "Sampl" in c]
print("sample columns:", cands)
sc = ci[[c for c in cands if "All" in c][0] if any("All" in c for c in cands) else cands[0]]
stall_cols = [c for c in hdr if c.startswith("stall_") or "Stall" in c]
agg = []
for k, r in enumerate(rows):
    try: n = float(r[sc])
    except Exception: continue
    agg.append((n, k, r))
tot = sum(a[0] for a in agg) or 1
print("total samples", tot, "instructions", len(agg))
# cumulative by coarse region: find first/last DMMA index
dm = [k for n, k, r in agg if "DMMA" in r[ci["Source"]]]
if dm:
    lo, hi2 = min(dm), max(dm)
    pre = sum(n for n, k, r in agg if k < lo); mid = sum(n for n, k, r in agg if lo <= k <= hi2); post = sum(n for n, k, r in agg if k > hi2)
    print(f"samples before first DMMA {100*pre/tot:.1f}%  within DMMA span {100*mid/tot:.1f}%  after last DMMA {100*post/tot:.1f}%")
    print("--- top instructions AFTER the last DMMA (epilogue / fix-up)")
    for n, k, r in sorted([a for a in agg if a[1] > hi2], reverse=True)[:30]:
        print(f"{100*n/tot:6.2f}%  #{k:5d}  {r[ci['Source']].strip()[:110]}")
    print("--- top instructions overall")
for n, k, r in sorted(agg, reverse=True)[:top]:
    print(f"{100*n/tot:6.2f}%  #{k:5d}  {r[ci['Source']].strip()[:110]}")
