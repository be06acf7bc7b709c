"""Turn the scratch outputs of one GPU round (gpurun_out/<TAG>_*) into the committed, judged artifacts under profiles/:

  profiles/<OUT>_bench_lines_1gpu.jsonl        the JSON line of every bench run of the round
  profiles/<OUT>_sweep_full_<workload>.txt      per-launch summary of the `ncu --set full` capture (key raw metrics,
                                                executed SASS by opcode, basic blocks with stall reasons)
  profiles/<OUT>_launches_<workload>.csv        the ncu launch list (gpu__time_duration.sum), if the round made one
  profiles/ncu_traffic.json                     dram bytes per sweep_kernel launch (read by bench.py: roofline.traffic)

Usage: python tools/make_profiles.py TAG [OUT]      (OUT defaults to TAG)"""
import csv, glob, io, json, os, subprocess, sys, contextlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else tag
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "lts__t_sector_hit_rate.pct"]

lines = []
for f in sorted(glob.glob(os.path.join(G, f"{tag}_bench*.json"))):
    for l in open(f):
        l = l.strip()
        if l.startswith("{"):
            lines.append(l)
if lines:
    with open(os.path.join(P, f"{out}_bench_lines_1gpu.jsonl"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("bench lines:", len(lines))

traffic_path = os.path.join(P, "ncu_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for rep in sorted(glob.glob(os.path.join(G, f"{tag}_sweep_full_*.ncu-rep"))):
    wl = os.path.basename(rep)[len(tag) + len("_sweep_full_"):-len(".ncu-rep")]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    txt = [f"# ncu --set full --clock-control none --import-source on -k regex:<sweep kernel> -s 3 -c 3 python bench.py --workload {wl} --steps 1 --warmup 1 --no-cpu-baseline --no-extra",
           f"# ({tag}); {len(data)} consecutive sweep_kernel launches of {wl}; numbers are per launch (cold caches, serialised replays)"]
    stall = [k for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k]
    tot = []
    for li, r in enumerate(data):
        txt.append("-----")
        txt.append(f"{'Kernel Name':92s}{r[hdr.index('Kernel Name')]}")
        for k in KEYS + stall:
            if k in hdr:
                i = hdr.index(k)
                txt.append(f"{k:92s}{r[i]:>20s} {units[i]}")
        rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[hdr.index("dram__bytes_read.sum")]]
        tot.append((rd + wr) * scale)
    traffic[wl] = int(round(sum(tot) / len(tot), -6))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", "0",
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    tmp = f"/tmp/_{tag}_{wl}_src.csv"
    open(tmp, "w").write(src)
    for tool, extra in (("ncu_sass_summary.py", ["14"]), ("ncu_blocks.py", ["1.5"])):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), tmp] + extra, capture_output=True, text=True)
        txt.append(f"----- first launch, {tool} (executed SASS of the capture's source page)")
        txt.append(r.stdout.rstrip())
    with open(os.path.join(P, f"{out}_sweep_full_{wl}.txt"), "w") as f:
        f.write("\n".join(txt) + "\n")
    print("ncu summary:", wl, "traffic/launch", traffic[wl])
json.dump(traffic, open(traffic_path, "w"), indent=1)
# ---- profiles/README.md: index + table of the round's bench lines
def row(d):
    r, e = d.get("roofline") or {}, d.get("e2e") or {}
    cfg = d.get("config", {})
    if d.get("impl") == "reference":
        return f"| {cfg.get('workload')} (reference `dmsim_cpu_omp`, {d['cpu_baseline']['cores']} host threads) | {d['n_gpus']} | – | {d['ms_per_step']:.0f} | {d['value']:.2f} | – | – | – | – | {d['cpu_baseline']['sample']} |"
    res = (e.get("resident_state") or {}).get("ms_per_step")
    jit = d.get("jit") or {}
    ach = f"{r.get('bound', 'hbm')}: {r.get('achieved', 0):.1f} {r.get('unit', 'GB/s')} ({r.get('frac', 0):.3f})"
    return (f"| {cfg.get('workload')} | {d['n_gpus']} | {cfg.get('sweeps_per_step')} | {d['ms_per_step']:.2f} | {d['value']:.0f} | "
            f"{ach} | {jit.get('sweeps_specialised', 0)}/{jit.get('sweeps_per_step', cfg.get('sweeps_per_step'))} | {e.get('ms_per_step', 0):.2f} ({e.get('value', 0):.0f}) | "
            f"{'' if res is None else format(res, '.2f')} | {(d.get('comm') or {}).get('GBps_per_direction') or ''} |")

md = ["# Bench lines under profiles/ (generated by tools/make_profiles.py; the narrative is in README.md)", "",
      "`ms/step` = device time of `dmb_run` (CUDA events on the engine's stream), sparse start off; `achieved` is GB/s when the",
      "binding floor is HBM and FMA-equivalent TFLOP/s when it is the FP64 pipe (`roofline.bound`); `e2e` = wall clock of reset +",
      "circuit upload + run + diagonal readback from |0><0|; `jit` = sweeps of a step that ran on run-time specialised kernels.", "",
      "| workload | GPUs | sweeps/step | ms/step | gates/s | bound: achieved (frac of the measured peak) | jit | e2e ms (gates/s) | resident e2e ms | exchange GB/s per direction | file |",
      "|---|---|---|---|---|---|---|---|---|---|---|"]
seen = set()
for f in sorted(glob.glob(os.path.join(P, "*bench_lines*.jsonl"))):
    if "older_kernel" in f:
        continue
    for l in open(f):
        l = l.strip()
        if l.startswith("{"):
            d = json.loads(l)
            key = (d["config"]["workload"], d["n_gpus"], d.get("impl"), os.path.basename(f))
            if key not in seen:
                seen.add(key)
                md.append(row(d) + f" `{os.path.basename(f)}` |")
open(os.path.join(P, "TABLE.md"), "w").write("\n".join(md) + "\n")
print("TABLE.md rows:", len(seen))
for f in glob.glob(os.path.join(G, f"{tag}_launches_*.csv")):
    dst = os.path.join(P, os.path.basename(f).replace(tag, out, 1))
    open(dst, "w").write(open(f).read())
    print("launch list:", os.path.basename(dst))
