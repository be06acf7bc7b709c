# round 2, multi-GPU call: both multi-GPU forms against the oracle (incl. the specialised kernels), then the default bench line of N GPUs.
# Usage: tools/gpu_multi_r2.sh TAG N
TAG=${1:-r2m}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt
(time timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs) > gpurun_out/${TAG}_pytest_multi.log 2>&1
tail -8 gpurun_out/${TAG}_pytest_multi.log
(time timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3) > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -c 1500 gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_bench_${N}gpu.json") if l.startswith("{")][-1])
    print("ms/step %.2f"%d["ms_per_step"], d["roofline"]["bound"], "frac %.3f"%d["roofline"]["frac"], "e2e %.2f"%d["e2e"]["ms_per_step"], d["jit"])
    print("comm", d.get("comm")); print("parity", json.dumps(d.get("parity"))[:900]); print("strong", d.get("strong_scaling"))
except Exception as e:
    print("FAILED", e)
PY
