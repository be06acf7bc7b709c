mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_multi.py -m gpu -x -q) > gpurun_out/r1k_pytest_multi4.log 2>&1
grep -E "passed|failed|error" gpurun_out/r1k_pytest_multi4.log | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r1k_bench4_default.json 2> gpurun_out/r1k_bench4_default.err
tail -c 900 gpurun_out/r1k_bench4_default.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 3 --warmup 3 --workload qft_n16 > gpurun_out/r1k_bench4_qft_n16.json 2> gpurun_out/r1k_bench4_qft_n16.err
tail -c 500 gpurun_out/r1k_bench4_qft_n16.json; echo
