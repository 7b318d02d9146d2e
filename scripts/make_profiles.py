"""Turn the raw gpurun_out/ captures of scripts/gpu_check.sh (NCU=1), diag_timeline.py and bench.py into the tracked
summaries under profiles/.  usage: python scripts/make_profiles.py [tag]"""
import collections, csv, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
# ---- launch list -> shares
rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, data = r, rows[i + 1:]
        break
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in data:
    if len(r) > mv and r[mn] == "gpu__time_duration.sum":
        agg.setdefault(r[kn].split("(")[0].replace("qmpc::", "").replace("void ", ""), []).append(float(r[mv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(P, f"{tag}_launch_shares.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -s 270 -c 90, closed-loop steps 30..39 of scripts/profile_step.py\n"
            "(B=4096, N=20, M=20, fp64, single stream). Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n")
    for k, v in agg.items():
        f.write(f"{k:50s} n={len(v):3d} mean={sum(v) / len(v) / 1000:9.1f} us  share={100 * sum(v) / tot:5.1f}%\n")
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_launches_ncu.csv"))
# ---- ncu --set full summaries of the two solver launches
with open(os.path.join(P, f"{tag}_ipm_summary.txt"), "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k 'regex:qmpc_ipm|qmpc_dense' -s 80 -c 2, closed-loop step 40 of scripts/profile_step.py\n"
            "(B=4096, N=20, M=20, fp64, single stream; default solver = Riccati screening launch + dense launch)\n")
    for title, rx in (("screening launch", "qmpc_ipm"), ("dense launch", "qmpc_dense")):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), os.path.join(G, "prof_ipm.ncu-rep"), rx],
                             capture_output=True, text=True).stdout
        f.write(f"\n=== {title}\n{out}")
# ---- timeline
with open(os.path.join(P, f"{tag}_timeline.txt"), "w") as f:
    f.write("Per-vehicle timeline of the solver launch(es) of ONE control step (qmpc_timeline_*: every OCP stamps %globaltimer when its\n"
            "warp/CTA starts and when its result is written), scripts/diag_timeline.py, B=4096 N=20 M=20 fp64, single stream, steps 40..42.\n"
            "'in flight' = OCPs running at 0 %, 5 %, ... 100 % of the span.  it = IPM iterations, rd = active-set rounds.\n\n"
            "################ A. Riccati kernel alone (QMPC_IPM_VARIANT=0; the kernel before the screening/dense split)\n"
            "# ~95 % of the vehicles finish within 0.2-0.8 ms (warm-started rounds); the span is set by the ~5 % that need the\n"
            "# interior-point method: 1-2.5 ms of ONE warp each, while the GPU idles (about 200 OCPs on 148 SMs after 25 % of the span).\n")
    f.write(open(os.path.join(G, "timeline_v1.txt")).read())
    f.write("\n################ B. default: Riccati screening launch (<= 6 warm-started rounds) + dense launch for the OCPs that did not settle\n"
            "# the dense launch starts when the screening launch has drained (start_us of the it>0 rows); a hard OCP now takes 0.2-0.5 ms on a\n"
            "# 256-thread CTA instead of 1-2.5 ms on one warp.\n")
    f.write(open(os.path.join(G, "timeline_now.txt")).read())
for src, dst in (("bench_N1.json", f"{tag}_bench_N1.json"), ("bench_N8.json", f"{tag}_bench_N8.json"), ("ubench_lat.txt", f"{tag}_ubench_latency.txt")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
print(open(os.path.join(P, f"{tag}_launch_shares.txt")).read())
