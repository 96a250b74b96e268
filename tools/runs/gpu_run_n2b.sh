set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
(timeout 600 python -m pytest tests/test_multi_device_gpu.py tests/test_p2p_mix.py tests/test_dispatch_gpu.py -x -q -m gpu) > gpurun_out/r3j_pytest_n2.log 2>&1
tail -4 gpurun_out/r3j_pytest_n2.log
timeout 600 python tools/multi_gpu_check.py --mode inproc --gpus 2 > gpurun_out/r3j_check_inproc_n2.jsonl 2> gpurun_out/r3j_check_inproc_n2.err
cut -c1-400 gpurun_out/r3j_check_inproc_n2.jsonl; tail -2 gpurun_out/r3j_check_inproc_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_check.py --mode ranks > gpurun_out/r3j_check_ranks_n2.jsonl 2> gpurun_out/r3j_check_ranks_n2.err
cut -c1-400 gpurun_out/r3j_check_ranks_n2.jsonl; tail -2 gpurun_out/r3j_check_ranks_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r3j_bench_n2.json 2> gpurun_out/r3j_bench_n2.err
cut -c1-300 gpurun_out/r3j_bench_n2.json; tail -2 gpurun_out/r3j_bench_n2.err
timeout 300 build/bin/cfg4_multi 2 3 > gpurun_out/r3j_cfg4_c_n2.json 2>&1; cat gpurun_out/r3j_cfg4_c_n2.json
