set -x
mkdir -p gpurun_out
nvidia-smi -L
(timeout 600 python -m pytest tests/test_multi_device_gpu.py tests/test_p2p_mix.py -x -q -m gpu) > gpurun_out/r2c_pytest_n2.log 2>&1
tail -15 gpurun_out/r2c_pytest_n2.log
timeout 600 python tools/multi_gpu_check.py --mode inproc --gpus 2 > gpurun_out/r2c_check_inproc_n2.jsonl 2> gpurun_out/r2c_check_inproc_n2.err
cat gpurun_out/r2c_check_inproc_n2.jsonl; tail -3 gpurun_out/r2c_check_inproc_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_check.py --mode ranks > gpurun_out/r2c_check_ranks_n2.jsonl 2> gpurun_out/r2c_check_ranks_n2.err
cat gpurun_out/r2c_check_ranks_n2.jsonl; tail -3 gpurun_out/r2c_check_ranks_n2.err
timeout 300 build/bin/cfg4_multi 2 3 > gpurun_out/r2c_cfg4_c_n2.json 2>&1; cat gpurun_out/r2c_cfg4_c_n2.json
timeout 300 build/bin/cfg4_multi 1 3 > gpurun_out/r2c_cfg4_c_n1.json 2>&1; cat gpurun_out/r2c_cfg4_c_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
cut -c1-400 gpurun_out/r2c_bench_n2.json; tail -3 gpurun_out/r2c_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/scale_timeline.py > gpurun_out/r2c_timeline_n2.json 2> gpurun_out/r2c_timeline_n2.err
cat gpurun_out/r2c_timeline_n2.json; tail -3 gpurun_out/r2c_timeline_n2.err
