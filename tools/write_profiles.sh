#!/bin/bash
# Turns the captures tools/gpu_profile.sh left in gpurun_out/ into the tracked text summaries under profiles/.
{
echo "# ncu --set full --clock-control none --import-source on captures (one launch each, B = 64; tools/gpu_profile.sh, tools/ncu_summary.py)"
echo "# times under ncu are cold-cache and serialised; bench.py's live per-launch event timings are the ones quoted in DESIGN.md"
for r in tma32_conv1 tma32_conv2 tma2_128_conv2 up64 tma2_sc32 attn; do
  echo; echo "== $r"
  case $r in
    tma32_conv1) echo "# ResBlock conv1, 32->32 @256x256 (conv_tma.cu): fp32 stream in (537 MB) -> fp16 h out (268 MB); algorithmic 805 MB";;
    tma32_conv2) echo "# ResBlock conv2, 32->32 @256x256 (conv_tma.cu): fp16 h in (268 MB) + fp32 residual (537 MB) -> fp32 stream (537 MB); algorithmic 1342 MB";;
    tma2_128_conv2) echo "# ResBlock conv2, 128->128 @64x64 (conv_tma2.cu): 77.3 GFLOP; fp16 in 67 MB + fp32 residual 134 MB -> fp32 out 134 MB";;
    up64) echo "# upsample+conv 64ch 128x128 -> 256x256 (conv_up.cu): fp16 in 134 MB -> fp32 out 1074 MB + fp16 copy 537 MB; algorithmic 1745 MB";;
    tma2_sc32) echo "# conv2 32->32 @256x256 with the fused 1x1 shortcut from 64 channels (conv_tma2.cu): fp16 h 268 MB + fp16 x 537 MB -> fp32 out 537 MB; algorithmic 1342 MB";;
    attn) echo "# attention core, B = 64, L = 1024, d = 128 (attention.cu): 34.4 GFLOP, 67 MB";;
  esac
  python tools/ncu_summary.py gpurun_out/prof_$r.ncu-rep
done
} > profiles/r1c_ncu_kernels.txt
python - <<'PY'
import csv, collections, re
with open('gpurun_out/launches.csv') as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    k = r['Kernel Name']; v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
    k = re.sub(r'\(.*$', '', k)[:72]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
out = ["# ncu launch list of `python bench.py --steps 2 --warmup 1` (first 900 launches; gpu__time_duration.sum, --clock-control none)",
       "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "",
       "%-74s %8s %12s %7s" % ("kernel", "launches", "total_us", "share")]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-74s %8d %12.1f %6.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
open('profiles/r1c_launch_list_summary.txt', 'w').write("\n".join(out) + "\n")
PY
