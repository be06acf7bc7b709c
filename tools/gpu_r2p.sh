TAG=${1:-r2p}
mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra"
for tb in 12 10 9 8 7; do
  DMB_TILE_BITS=$tb $B --workload vqe_uccsd_n8 > gpurun_out/${TAG}_bench_tb${tb}_vqe_uccsd_n8.json 2> gpurun_out/${TAG}_bench_tb${tb}_vqe_uccsd_n8.err
  DMB_TILE_BITS=$tb $B --workload adder_n10 > gpurun_out/${TAG}_bench_tb${tb}_adder_n10.json 2> gpurun_out/${TAG}_bench_tb${tb}_adder_n10.err
done
DMB_TILE_BITS=9 DMB_PERSISTENT=0 $B --workload vqe_uccsd_n8 > gpurun_out/${TAG}_bench_tb9np_vqe_uccsd_n8.json 2> gpurun_out/${TAG}_bench_tb9np_vqe_uccsd_n8.err
for w in qft_n15 random_c1c2_n15; do
DMB_JIT=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dmb_jit_sweep -s 3 -c 3 -f -o gpurun_out/${TAG}_sweep_full_$w python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_full_$w.log 2>&1
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("${TAG}_bench_")[1][:-5], "ms/step %.3f"%d["ms_per_step"], "sweeps", d["config"]["sweeps_per_step"], "launches", d["gpu_launches"], "e2e %.2f"%d["e2e"]["ms_per_step"], "warm %.2f"%d["e2e"]["repeated_circuit"]["ms_per_step"], "trace %.15f"%d["trace_after_run"])
    except Exception as e:
        print(f, "FAILED", e)
PY
