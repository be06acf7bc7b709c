# round 2, call B (1 GPU): TMA with conflict-aware groups, A/B + ncu.  Usage: tools/gpu_r2b.sh TAG
TAG=${1:-r2b}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
B="timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
for w in qft_n15 bv_n15 hlayer_n15 random_c1c2_n15; do
  DMB_TMA=1 $B --workload $w > gpurun_out/${TAG}_bench_tma_$w.json 2> gpurun_out/${TAG}_bench_tma_$w.err
done
DMB_TMA=1 DMB_TMA_BOX_BITS=12 $B --workload qft_n15 > gpurun_out/${TAG}_bench_tma_box12_qft_n15.json 2> gpurun_out/${TAG}_bench_tma_box12_qft_n15.err
for W in qft_n15 random_c1c2_n15; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 3 -f -o gpurun_out/${TAG}_sweep_full_$W python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_$W.log 2>&1
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("${TAG}_bench_")[1][:-5], "ms/step %.2f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.2f"%d["e2e"]["ms_per_step"], "trace %.15f"%d["trace_after_run"])
    except Exception as e:
        print(f, "FAILED", e)
PY
