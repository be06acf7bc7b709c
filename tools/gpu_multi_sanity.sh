# 2-GPU sanity: sharded parity test (peer-memory agreement path) + one bench line.  Usage: tools/gpu_multi_sanity.sh TAG
TAG=${1:-rX}
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_multi.py -m gpu -x -q) > gpurun_out/${TAG}_pytest_multi2.log 2>&1
grep -E "passed|failed|error" gpurun_out/${TAG}_pytest_multi2.log | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --workload qft_n16 > gpurun_out/${TAG}_bench2_qft_n16.json 2> gpurun_out/${TAG}_bench2_qft_n16.err
tail -c 400 gpurun_out/${TAG}_bench2_qft_n16.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench2_random_c1c2_n16.json 2> gpurun_out/${TAG}_bench2_random_c1c2_n16.err
tail -c 400 gpurun_out/${TAG}_bench2_random_c1c2_n16.json; echo
