set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r3k_bench_n8.json 2> gpurun_out/r3k_bench_n8.err
cut -c1-330 gpurun_out/r3k_bench_n8.json; echo; tail -2 gpurun_out/r3k_bench_n8.err
timeout 200 build/bin/cfg4_multi 8 5 > gpurun_out/r3k_cfg4_c_n8.json 2>&1; cat gpurun_out/r3k_cfg4_c_n8.json
CFG5_WAV=/tmp/cfg5.wav timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 tools/render_cfg5.py > gpurun_out/r3k_cfg5_n8.json 2> gpurun_out/r3k_cfg5_n8.err
cut -c1-900 gpurun_out/r3k_cfg5_n8.json; tail -2 gpurun_out/r3k_cfg5_n8.err
