"""Turn gpurun_out/ ncu artefacts into small text/JSON summaries under profiles/ (tracked).
    python scripts/summarize_profiles.py <tag>     e.g. r01_v5
Reads (when present): gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum launch list),
gpurun_out/prof_omc.ncu-rep, gpurun_out/prof_ret.ncu-rep (ncu --set full), gpurun_out/bench.json."""
import csv, collections, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
G = os.path.join(ROOT, "gpurun_out")


def launch_list():
    p = os.path.join(G, "launches.csv")
    if not os.path.exists(p):
        return
    rows = list(csv.reader(open(p)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols = rows[h]
    ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki], [0.0, 0])
        a[0] += v
        a[1] += 1
    tot = sum(a[0] for a in agg.values())
    with open(os.path.join(out_dir, f"{tag}_ncu_launch_list.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# command: python bench.py --steps 20 --warmup 3 --no-cpu   (launches after the first 60)\n")
        f.write(f"# {'avg_us':>9s} {'launches':>8s} {'share':>6s}  kernel\n")
        for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"  {t / c:9.2f} {c:8d} {100 * t / tot:5.1f}%  {k[:150]}\n")


def full_report(rep, name, keys):
    p = os.path.join(G, rep)
    if not os.path.exists(p):
        return
    raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(out_dir, f"{tag}_{name}_ncu_full.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({rep}); selected raw metrics per profiled launch\n")
        for r in rows[2:]:
            f.write(f"== {r[idx['Kernel Name']][:140]}\n")
            for k in keys:
                for hname in hdr:
                    if hname == k or hname.endswith("." + k) or hname.endswith(k):
                        if r[idx[hname]]:
                            f.write(f"   {hname:95s} {r[idx[hname]]:>18s} {units[idx[hname]]}\n")
                        break
    top = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_top.py"), p, "25"], capture_output=True, text=True).stdout
    with open(os.path.join(out_dir, f"{tag}_{name}_ncu_source_top.txt"), "w") as f:
        f.write("# top stall-sampled SASS instructions (ncu --page source), per profiled launch\n" + top)


KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "smsp__issue_inst0.avg.pct_of_peak_sustained_active"]
launch_list()
full_report("prof_omc.ncu-rep", "omc_gemms", KEYS)
full_report("prof_ret.ncu-rep", "sim_topk", KEYS)
b = os.path.join(G, "bench.json")
if os.path.exists(b):
    lines = [l for l in open(b).read().splitlines() if l.startswith("{")]
    if lines:
        json.dump(json.loads(lines[-1]), open(os.path.join(out_dir, f"{tag}_bench.json"), "w"), indent=1)
print(sorted(os.listdir(out_dir)))
