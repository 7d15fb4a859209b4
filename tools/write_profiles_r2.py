"""Turns what tools/gpu_profile_r2.sh left in gpurun_out/ into the tracked round-2 evidence under profiles/:
ncu summaries, launch lists, bench / training JSON lines, single-op timings, DRAM traffic per kernel class and the
per-kernel SASS census of libptivae.so (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = TMEM)."""
import collections
import csv
import io
import json
import pathlib
import re
import shutil
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
P.mkdir(exist_ok=True)

CAPS = [
    ("band32c2", "ResBlock conv2 32->32 @256x256, B = 64, fp16 stream (conv_band.cu <32,2>): in 268 MB + residual 268 MB -> out 268 MB; SURVEY 8(d) bytes 537 MB (+268 residual); 77.3 GFLOP"),
    ("band32c1", "ResBlock conv1 32->32 @256x256, B = 64, fp16 stream (conv_band.cu <32,0>): in 268 MB -> out 268 MB = SURVEY 8(d) bytes 537 MB; 77.3 GFLOP"),
    ("gnstats32", "gn_stats, [64,256,256,32] fp16 (gn.cu): 268 MB read"),
    ("gnapply32", "gn_apply + SiLU, [64,256,256,32] fp16 -> fp16 (gn.cu): 268 MB read + 268 MB written"),
    ("l1l2", "l1l2_partial, 2 x [64,1,256,256] fp32 (latent_loss.cu): 33.6 MB read"),
    ("cin1", "conv3x3_cin1 1->32 @256x256 + first-norm statistics, fp32 out (conv_direct.cu): 16.8 MB in -> 537 MB out"),
    ("cout1", "conv3x3_fewcout 32->1 @256x256 with the final GroupNorm affine folded in (conv_direct.cu): 268 MB fp16 in -> 16.8 MB out"),
    ("s128c2", "ResBlock conv2 128->128 @64x64, B = 64, fp16 stream, ONE-SM kernel (conv_tma2.cu <128,128,0,2,0,0>, PTIVAE_PAIR=0): 77.3 GFLOP; in 67 MB + residual 67 MB -> out 67 MB.  Shared-memory data pipe: tensor-core reads (tc) + LSU (LDS / STS / TMA) wavefronts"),
    ("p128c2", "the same layer on the TWO-SM kernel (conv_pair.cu <128,128,2,2>, cta_group::2): each SM stages half of the weight rows"),
    ("p128c1", "ResBlock conv1 128->128 @64x64 on the two-SM kernel (no residual)"),
    ("p256c1", "config B: conv1 256->256 @64x64, B = 64, two-SM kernel (conv_pair.cu <256,256,0,2>): 309 GFLOP"),
    ("s256c1", "config B: conv1 256->256 @64x64, B = 64, one-SM kernel (conv_tma2.cu <256,256,0,0,0,0>, captured before the packed-half2 prologue)"),
    ("s64c2", "ResBlock conv2 64->64 @128x128, B = 64, fp16 stream (conv_tma2.cu <64,64,0,2,0,0>, four epilogue teams): 77.3 GFLOP; in 134 MB + residual 134 MB -> out 134 MB"),
    ("s128c2s", "ResBlock conv2 128->128 @32x32, B = 64, fp16 stream, one-SM kernel (captured before the packed-half2 prologue): 256 tiles on 148 CTAs"),
    ("wgrad_col", "wgrad3x3_col_kernel (wgrad.cu), one launch of an eager B = 8 training step"),
    ("gn_bwd_reduce", "gn_bwd_reduce_kernel (gn_bwd.cu), one launch of an eager B = 8 training step"),
    ("dgrad_umma", "conv_umma_kernel as a data-gradient conv (conv_umma.cu mode 4), one launch of an eager B = 8 training step"),
]


def ncu_raw(rep):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[1], rows[2]


def main():
    out = ["# ncu --set full --clock-control none --import-source on, one launch each (tools/gpu_profile_r2.sh, tools/ncu_summary.py)",
           "# times under ncu are cold-cache and serialised; the event timings of tools/bench_ops.py (profiles/r2_ops.txt) and of bench.py are the ones quoted"]
    traffic = json.loads((P / "traffic.json").read_text()) if (P / "traffic.json").exists() else {}
    for name, desc in CAPS:
        rep = G / f"r2prof_{name}.ncu-rep"
        if not rep.exists():
            continue
        out += ["", f"== {name}", f"# {desc}"]
        out.append(subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(rep)], capture_output=True, text=True).stdout.rstrip())
        hdr, units, vals = ncu_raw(rep)

        def val(k):
            if k not in hdr:
                return None
            v, u = float(vals[hdr.index(k)].replace(",", "")), units[hdr.index(k)]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            key = {"band32c2": "fused3x3_32->32@256x256_in2_res2_out2", "band32c1": "fused3x3_32->32@256x256_in2_res0_out2"}.get(name)
            if key:
                def pct(k):
                    return float(vals[hdr.index(k)].replace(",", "")) if k in hdr else None
                tc = pct("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")
                lsu = pct("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")
                traffic[key] = {"dram_bytes_per_launch": rd + wr,
                                "note": f"ncu --set full: {rd / 1e6:.1f} MB read + {wr / 1e6:.1f} MB write",
                                "source": f"profiles/r2_ncu_kernels.txt ({name})"}
                if tc is not None and lsu is not None:   # what actually limits the kernel: the shared-memory data pipe
                    traffic[key]["shared_memory_pipe_pct_of_peak"] = {"tensor_core_operand_reads": tc, "lsu_lds_sts_tma": lsu}
    (P / "r2_ncu_kernels.txt").write_text("\n".join(out) + "\n")
    (P / "traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")

    # inference launch list
    lst = G / "r2_launches.csv"
    if lst.exists():
        lines = [l for l in lst.read_text().splitlines() if not l.startswith("==")]
        agg = collections.OrderedDict()
        for r in csv.DictReader(lines):
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            k, v, u = r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
            v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
            k = re.sub(r"\(.*$", "", k)[:80]
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(a[1] for a in agg.values())
        o = ["# ncu launch list of `python bench.py --steps 2 --warmup 1 --no-eager-baseline` (first 900 launches; gpu__time_duration.sum, --clock-control none)",
             "# per-launch times are cold-cache and serialised: compare SHARES with bench.py's live breakdown, not absolutes", "",
             "%-82s %8s %12s %7s" % ("kernel", "launches", "total_us", "share")]
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.append("%-82s %8d %12.1f %6.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
        (P / "r2_launch_list_summary.txt").write_text("\n".join(o) + "\n")

    for src, dst in (("r2_bench.json", "r2_bench.json"), ("r2_bench_ref.json", "r2_bench_ref.json"), ("r2_train_b8.json", "r2_train_b8.json"),
                     ("r2_train_b32.json", "r2_train_b32.json"), ("r2_train_launch_summary_b8.txt", "r2_train_launch_summary_b8.txt"),
                     ("r2_ops.log", "r2_ops.txt"), ("r2_ops_wide.log", "r2_ops_wide.txt"), ("r2_timeline_wide.log", "r2_timeline_wide.txt"), ("r2_timeline_band.log", "r2_timeline_band.txt"),
                     ("r2_sfu_rates.log", "r2_sfu_rates.txt"), ("r2_train_b8_after.json", "r2_train_b8_after.json"), ("bench_breakdown.json", "r2_bench_breakdown.json"),
                     ("r2_train_full.json", "r2_train_full.json"), ("r2_configs.json", "r2_configs.json"),
                     ("r2_bench_n2.json", "r2_bench_n2.json"), ("r2_train_b8_n2.json", "r2_train_b8_n2.json"),
                     ("r2_train_full_n2.json", "r2_train_full_n2.json"), ("r2_bench_n4.json", "r2_bench_n4.json"),
                     ("r2_train_b8_n4.json", "r2_train_b8_n4.json"), ("r2_train_full_n4.json", "r2_train_full_n4.json")):
        if (G / src).exists():
            # keep the explanatory header ('#' lines) of an already committed record when its raw log is copied again
            head = ""
            if dst.endswith(".txt") and (P / dst).exists():
                old = (P / dst).read_text().splitlines(keepends=True)
                head = "".join(l for l in old[:12] if l.startswith("#"))
            new = (G / src).read_text() if dst.endswith(".txt") else None
            if new is not None and head and not new.startswith("#"):
                (P / dst).write_text(head + new)
            else:
                shutil.copy(G / src, P / dst)

    # SASS census of the shipped library
    so = ROOT / "pti-ldm-vae_b200" / "libptivae.so"
    if so.exists():
        sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
        cur, cnt = None, collections.OrderedDict()
        keys = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "HMMA", "MUFU")
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                cur = re.sub(r"\(.*$", "", name).replace("ptivae::", "")[:110]
                cnt[cur] = collections.Counter()
                continue
            if cur is None:
                continue
            for k in keys:
                if re.search(r"\b" + k + r"\b|\b" + k + r"\.", line):
                    cnt[cur][k] += 1
        o = ["# SASS census of pti-ldm-vae_b200/libptivae.so (cuobjdump -sass; sm_100a only): static instruction counts per kernel",
             "# UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit",
             "# (no HMMA anywhere: nothing runs on the legacy mma.sync path)", "",
             "%-112s " % "kernel" + " ".join("%8s" % k for k in keys)]
        tot = collections.Counter()
        for k, c in cnt.items():
            if sum(c.values()) == 0:
                continue
            o.append("%-112s " % k + " ".join("%8d" % c[x] for x in keys))
            tot.update(c)
        o.append("%-112s " % "TOTAL" + " ".join("%8d" % tot[x] for x in keys))
        (P / "r2_sass_summary.txt").write_text("\n".join(o) + "\n")
    print("profiles written:", sorted(p.name for p in P.iterdir() if p.name.startswith("r2_")))


if __name__ == "__main__":
    main()
