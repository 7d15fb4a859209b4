"""Samples per CUDA source line (ncu --print-source cuda,sass)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
cur_file = "?"; hdr = None; tot = 0.0; toti = 0.0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; ix = {}; [ix.setdefault(h, i) for i, h in enumerate(hdr)]; stalls = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr is None or len(r) != len(hdr): continue
    try:
        smp = float(r[ix["# Samples"]] or 0); ins = float(r[ix["Instructions Executed"]] or 0)
    except ValueError: continue
    key = (cur_file, r[0], r[1].strip()[:100])
    a = agg[key]; a[0] += smp; a[1] += ins
    for i, n in stalls:
        try: a[2][n] += float(r[i] or 0)
        except ValueError: pass
    tot += smp; toti += ins
print("total samples", tot, "instr", toti)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    st = ", ".join(f"{n}:{v:.0f}" for n, v in a[2].most_common(2))
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[1]/toti:5.1f}%  {key[0]}:{key[1]:>4s}  {key[2]:100s} [{st}]")
