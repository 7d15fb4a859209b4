"""Aggregate an ncu source page (--print-source cuda,sass not needed): samples per CUDA source line."""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# find header
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data); toti = sum(f(r, "Instructions Executed") for r in data)
print("total samples", tot, "warp instr", toti)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print("stalls:", ", ".join(f"{k[6:]}={100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
top = sorted(data, key=lambda r: -f(r, "# Samples"))[:int(sys.argv[2]) if len(sys.argv) > 2 else 45]
for r in top:
    st = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    src = r[ix["Source"]].strip()[:95]
    print(f"{100*f(r,'# Samples')/tot:5.1f}% inst={100*f(r,'Instructions Executed')/toti:5.1f}% {r[ix.get('Address', 0)][:8]:>8s} {src:95s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
