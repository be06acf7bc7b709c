# N-GPU box call: sharded parity test (peer-memory and NCCL exchange paths) + bench lines.  Usage: tools/gpu_multi.sh TAG N
TAG=${1:-rX}
N=${2:-2}
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_multi.py -m gpu -x -q) > gpurun_out/${TAG}_pytest_multi${N}.log 2>&1
grep -E "passed|failed|error" gpurun_out/${TAG}_pytest_multi${N}.log | tail -2
(DMB_P2P=0 python -m pytest tests/test_gpu_multi.py -m gpu -x -q) > gpurun_out/${TAG}_pytest_multi${N}_nccl.log 2>&1
grep -E "passed|failed|error" gpurun_out/${TAG}_pytest_multi${N}_nccl.log | tail -2
run() { # gpus workload steps tag-suffix
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $1 --steps $3 --warmup 3 --workload $2 > gpurun_out/${TAG}_bench${1}_$2$4.json 2> gpurun_out/${TAG}_bench${1}_$2$4.err
  tail -c 700 gpurun_out/${TAG}_bench${1}_$2$4.json; echo
}
if [ "$N" = "8" ]; then
  run 8 random_c1c2_n17 3
  run 8 qft_n17 3
  run 4 random_c1c2_n16 3
  run 8 qft_n15 5
else
  run $N random_c1c2_n16 3
  DMB_P2P=0 run $N random_c1c2_n16 3 _nccl
  run $N qft_n16 3
  DMB_P2P=0 run $N qft_n16 3 _nccl
fi
