set -x
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_multi_device_gpu.py -x -q -m gpu) > gpurun_out/r2l_pytest_multi.log 2>&1; tail -3 gpurun_out/r2l_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2l_bench_n8.json 2> gpurun_out/r2l_bench_n8.err
cut -c1-330 gpurun_out/r2l_bench_n8.json; echo; tail -2 gpurun_out/r2l_bench_n8.err | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 tools/scale_timeline.py > gpurun_out/r2l_timeline_n8.json 2> gpurun_out/r2l_timeline_n8.err
cut -c1-700 gpurun_out/r2l_timeline_n8.json; echo
timeout 200 build/bin/cfg4_multi 8 5 > gpurun_out/r2l_cfg4_c_n8.json 2>&1; cat gpurun_out/r2l_cfg4_c_n8.json
CFG5_WAV=/tmp/cfg5.wav timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 tools/render_cfg5.py > gpurun_out/r2l_cfg5_n8.json 2> gpurun_out/r2l_cfg5_n8.err
cat gpurun_out/r2l_cfg5_n8.json; tail -3 gpurun_out/r2l_cfg5_n8.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2l_bench_n4.json 2> gpurun_out/r2l_bench_n4.err
cut -c1-330 gpurun_out/r2l_bench_n4.json; echo
