import json,glob,sys
tag=sys.argv[1]
for f in sorted(glob.glob(f'gpurun_out/{tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f"{d['config']['workload']:18s} ms/step {d['ms_per_step']:8.2f} sweeps {d['config']['sweeps_per_step']:3d} GB/s {d['roofline']['achieved']:6.0f} frac {d['roofline']['frac']:.3f} gates/s {d['value']:.0f} e2e {d['e2e']['value']:.0f}")
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
