"""Summarise the ncu artefacts of one bench command into profiles/ (text summary + per-kernel DRAM traffic json).
usage: python tools/ncu_summary.py <launch_list.csv> <full_report.ncu-rep> <bench_default.json> <round tag, e.g. r01> [more .ncu-rep ...]

  launch list : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file <launch_list.csv> python bench.py ...
  full report : ncu --set full --clock-control none --import-source on -k regex:... -o <report> python bench.py ...
"""
import csv, io, json, os, re, subprocess, sys
from collections import OrderedDict

launch_csv, report, bench_json, tag = sys.argv[1:5]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:110]


# ---- launch list: per-kernel totals and shares
lines = [l for l in open(launch_csv) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = OrderedDict()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    k = short(r["Kernel Name"])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(v[1] for k, v in agg.items() if "k_sim" not in k and "k_gather" not in k)
lst = sorted(agg.items(), key=lambda kv: -kv[1][1])

# ---- full report: one row per captured launch
def read_report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(raw)))


rr = read_report(report)
hdr = rr[0]
col = {h: i for i, h in enumerate(hdr)}
want = [("time_us", "gpu__time_duration.sum"), ("dram_rd_GB", "dram__bytes_read.sum"), ("dram_wr_GB", "dram__bytes_write.sum"),
        ("regs", "launch__registers_per_thread"), ("warps_act%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        ("dmma%", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("l1tex%", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
        ("lts%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dram%", "dram__throughput.avg.pct_of_peak_sustained_elapsed")]
seen = OrderedDict()
all_rows = [(rr[1], col, r) for r in rr[2:]]
for extra in sys.argv[5:]:  # further reports of the same command (e.g. the solve kernels captured separately)
    er = read_report(extra)
    ecol = {h: i for i, h in enumerate(er[0])}
    all_rows += [(er[1], ecol, r) for r in er[2:]]
for units, col, r in all_rows:
    k = short(r[col["Kernel Name"]])
    if k in seen:
        continue
    d = {}
    for lab, m in want:
        if m in col and r[col[m]] != "":
            v = float(r[col[m]].replace(",", ""))
            u = units[col[m]]
            if lab == "time_us":
                v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
            if lab.endswith("_GB"):
                v = v / 1e9 if u in ("byte", "B") else v / 1e3 if u in ("Mbyte", "MB") else v / 1e6 if u in ("Kbyte", "KB") else v
            d[lab] = v
    # achieved fp64 arithmetic: thread-level DFMA (x2) + DMUL + DADD, and the tensor path's fp64 flops
    def val(m):
        return float(r[col[m]].replace(",", "")) if m in col and r[col[m]] not in ("", "n/a") else 0.0
    cyc = val("sm__cycles_elapsed.max")
    scalar = (2 * val("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") +
              val("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") +
              val("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed")) * cyc
    if d.get("time_us"):
        d["f64_TF/s"] = scalar / (d["time_us"] * 1e-6) / 1e12
        d["dmma_TF/s"] = val("sm__ops_path_tensor_src_fp64.sum") / (d["time_us"] * 1e-6) / 1e12
    seen[k] = d

bench = json.load(open(bench_json))
traffic = {}
for k, d in seen.items():
    if "dram_rd_GB" in d:
        traffic[k.split("::")[-1].split("<")[0]] = (d["dram_rd_GB"] + d.get("dram_wr_GB", 0.0)) * 1e9
wl = bench["config"]["workload"]
tj_path = os.path.join(out_dir, f"{tag}_traffic.json")
json.dump({wl: traffic}, open(tj_path, "w"), indent=1)

with open(os.path.join(out_dir, f"{tag}_ncu_summary.txt"), "w") as f:
    c = bench["config"]
    ws = bench.get("workload_stats", c)
    f.write(f"profiles/{tag} -- NVIDIA B200 (sm_100a), workload {wl}: {c['events']} events, {ws['measurements']} measurements, "
            f"{c['sensor'][0]}x{c['sensor'][1]} sensor, {c['panorama'][0]}x{c['panorama'][1]} panorama, n = {c['control_poses']} "
            f"control poses, Np = {ws['active_pixels']} active pixels\n\n")
    f.write("Commands (each ncu pass only after the same command exited 0 without ncu, same gpurun call):\n"
            "  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras\n"
            f"  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv   -> {tag}_launches_bench_{wl}.csv\n"
            "  ncu --set full --clock-control none --import-source on -k regex:\"k_eval|k_asm_pose|k_pix|k_place|k_seg_sort\" --launch-skip 14 --launch-count 7 (the third pass)\n"
            "  ncu --set full --clock-control none --import-source on -k regex:\"k_schur_tiles|k_ldlt_fused|k_solve_x2\" --launch-skip 3 --launch-count 3\n"
            f"Default bench of the same build (python bench.py): {tag}_bench_default_{wl}.json\n")
    rf = bench["roofline"]
    f.write(f"  value {bench['value']:.4g} events/s ({bench['ms_per_step']:.3f} ms per pass), e2e {bench['e2e']['value']:.4g} events/s, "
            f"clocks {bench['clocks']['sm_mhz']:.0f} MHz, throttle reasons {bench['clocks']['reasons']}\n")
    f.write(f"  roofline (live CUDA events, algorithmic bytes / duration): {rf['kernel']} {rf['achieved']:.0f} GB/s = "
            f"{rf['frac']:.3f} of {rf['peak']:.1f} GB/s ({rf['peak_source']})\n")
    f.write("  live kernel times (ms): " + ", ".join(f"{k} {v:.3f}" for k, v in rf["kernels_ms"].items()) + "\n")
    f.write("  algorithmic GB/s: " + ", ".join(f"{k} {v:.0f}" for k, v in rf["kernels_alg_gbs"].items()) + "\n")
    if "lm" in bench and bench["lm"] and "ms_per_iteration" in bench["lm"]:
        f.write(f"  LM: {bench['lm']['ms_per_iteration']:.3f} ms per iteration ({bench['lm'].get('solves', bench['lm'].get('iterations'))} solves, {bench['lm']['accepted']} accepted)\n")
    if "cpu_baseline" in bench:
        cb = bench["cpu_baseline"]
        f.write(f"  CPU {cb['kind']}: {cb['value']:.4g} {cb['unit']} on {cb['cores']} core(s)\n")
    f.write("\n== ncu --set full, one launch each (cold cache, serialised; shares matter, not absolutes) ==\n"
            "   (f64_TF/s: thread-level DFMA x2 + DMUL + DADD per second; dmma_TF/s: fp64 flops of the tensor path per second)\n")
    labs = [l for l, _ in want] + ["f64_TF/s", "dmma_TF/s"]
    f.write(f"{'kernel':<28}" + "".join(f"{l:>11}" for l in labs) + "\n")
    for k, d in seen.items():
        f.write(f"{k.split('::')[-1][:27]:<28}" + "".join(f"{d[l]:>11.3f}" if l in d else f"{'-':>11}" for l in labs) + "\n")
    f.write("\n== launch list of the bench command: time per kernel (synthetic-data simulator excluded from the shares) ==\n")
    for k, (cnt, us) in lst[:40]:
        share = "   sim" if ("k_sim" in k or "k_gather" in k) else f"{100 * us / total:5.1f}%"
        f.write(f"{us / cnt:10.1f} us/launch {cnt:5d}x {share}  {k}\n")
print("wrote", tj_path)
