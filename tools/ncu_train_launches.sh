#!/bin/bash
# Launch list (gpu__time_duration per kernel) of ONE eager training step at batch $1 (default 8): tools/bench_train.py runs
# 2 warm-up steps + 1 phase-split step + 3 warm-ups before the timed step; each step launches the same kernels, so
# skipping 3 steps' worth of launches and capturing one step's worth isolates a steady-state step.
B=${1:-8}
N=${2:-644}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((3 * N)) -c $N --csv \
    --log-file gpurun_out/train_launches_b$B.csv python tools/bench_train.py --batch $B --steps 1 --warmup 1 --no-graph \
    > gpurun_out/ncu_train.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/train_launches_b$B.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
tot = 0.0
for r in rows[hdr + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] in ("ns", "nsecond") else v      # -> us
    name = r[kn].split("(")[0].split("<")[0].replace("ptivae::", "")
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v
    a[1] += 1
    tot += v
print(f"total {tot:.1f} us over {sum(a[1] for a in agg.values())} launches")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{t:9.1f} us {n:4d}x {t / n:8.1f} us/launch  {100 * t / tot:5.1f}%  {k}")
PY
