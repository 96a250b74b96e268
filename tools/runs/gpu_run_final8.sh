set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2q_bench_n8.json 2> gpurun_out/r2q_bench_n8.err
cut -c1-330 gpurun_out/r2q_bench_n8.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 tools/scale_timeline.py 2> gpurun_out/r2q_timeline_n8.err | grep -v "^NCCL" > gpurun_out/r2q_timeline_n8.json
cut -c1-600 gpurun_out/r2q_timeline_n8.json; echo
timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2q_bench_n1.json 2>/dev/null; cut -c1-200 gpurun_out/r2q_bench_n1.json; echo
timeout 200 build/bin/cfg4_multi 8 5 > gpurun_out/r2q_cfg4_c_n8.json 2>&1; cat gpurun_out/r2q_cfg4_c_n8.json
