TAG=${1:-rX}
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1
for w in qft_n15 bv_n15 random_c1c2_n15 single_n15 hlayer_n15; do python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; done
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 4 -c 2 -f -o gpurun_out/${TAG}_sweep_full python bench.py --workload random_c1c2_n15 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(d['config']['workload'], 'ms/step %.2f'%d['ms_per_step'], 'sweeps', d['config']['sweeps_per_step'], 'GB/s %.0f'%d['roofline']['achieved'], 'frac %.3f'%d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
PY
