"""Turn gpurun_out/{train_launches.csv, prof_train_*.ncu-rep} into tracked summaries under profiles/ (run on the CPU box)."""
import collections, csv, io, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles"); SRC = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
lines = [l for l in open(os.path.join(SRC, "train_launches.csv")) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    try: v = float(row["Metric Value"].replace(",", ""))
    except ValueError: continue
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1.0)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    a = agg.setdefault(name, [0, 0.0, row["Grid Size"], row["Block Size"]]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(OUT, f"{tag}_train_launches_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none over `python scripts/train_probe.py --one`\n")
    f.write("# (ONE c3 training step: B=32 x 3 s, multitask, fwd+bwd+L2+clip+Adam; per-launch times are cold-cache and serialised: compare SHARES;\n")
    f.write("#  at::* rows are torch's own fills/copies/index ops of the host glue -- one_hot, zeros, masks)\n")
    f.write(f"{'kernel':72s} {'launches':>8s} {'total_ms':>10s} {'share':>7s}  grid block (last launch)\n")
    for k, (n, t, g, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k[:72]:72s} {n:8d} {t/1e6:10.3f} {t/tot:7.3f}  {g} {b}\n")
shutil.copy(os.path.join(SRC, "train_launches.csv"), os.path.join(OUT, f"{tag}_train_launches.csv"))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"]
for fn in sorted(os.listdir(SRC)):
    m = re.match(r"prof_train_(.*)\.ncu-rep", fn)
    if not m: continue
    k = m.group(1)
    raw = subprocess.run(["ncu", "-i", os.path.join(SRC, fn), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    if len(r) < 3: continue
    hdr, units, row = r[0], r[1], r[2]
    vals = {h: (v, u) for h, u, v in zip(hdr, units, row)}
    with open(os.path.join(OUT, f"{tag}_ncu_train_{k}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{k} -s <n> -c 1 python scripts/train_probe.py --one   (c3 training step)\n")
        f.write(f"# kernel: {vals.get('Kernel Name', ('?',))[0]}\n")
        for key in KEYS:
            for h in hdr:
                if h == key or h.endswith(key):
                    f.write(f"{h:90s} {vals[h][0]:>16s} {vals[h][1]}\n")
        stalls = sorted(((float(v[0]), h) for h, v in vals.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v[0] not in ("", "n/a")), reverse=True)[:8]
        f.write("# top warp-stall reasons (per issue-active):\n")
        for v, h in stalls: f.write(f"{h:90s} {v:16.3f}\n")
    print("wrote", k)
print(open(os.path.join(OUT, f"{tag}_train_launches_summary.txt")).read()[:2500])
