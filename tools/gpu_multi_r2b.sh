# Usage: tools/gpu_multi_r2b.sh TAG N : the specialised-kernel group test, then the default bench line of N GPUs
TAG=${1:-r2s}; N=${2:-2}
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || (time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "specialised") > gpurun_out/${TAG}_pytest_multi.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_multi.log
(time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3) > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -c 800 gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_bench_${N}gpu.json") if l.startswith("{")][-1])
    print("ms/step %.2f"%d["ms_per_step"], d["roofline"]["bound"], "frac %.3f"%d["roofline"]["frac"], "e2e %.2f"%d["e2e"]["ms_per_step"], "first", d["e2e"]["first_call_ms"], "cont", d["continued_state"]["ms_per_step"])
    print({k: d["jit"][k] for k in d["jit"] if k != "what"})
    print("comm", d.get("comm")); print("parity", d["parity"]["max_abs_err"], d["parity"]["ok"]); print("strong", d.get("strong_scaling"))
except Exception as e:
    print("FAILED", e)
PY
