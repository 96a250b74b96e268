set -x
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err
cut -c1-330 gpurun_out/r2f_bench_n$N.json; tail -2 gpurun_out/r2f_bench_n$N.err | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tools/scale_timeline.py > gpurun_out/r2f_timeline_n$N.json 2> gpurun_out/r2f_timeline_n$N.err
cut -c1-1500 gpurun_out/r2f_timeline_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 tools/multi_gpu_check.py --mode ranks > gpurun_out/r2f_check_ranks_n$N.jsonl 2> gpurun_out/r2f_check_ranks_n$N.err
cat gpurun_out/r2f_check_ranks_n$N.jsonl
timeout 600 python tools/multi_gpu_check.py --mode inproc --gpus $N > gpurun_out/r2f_check_inproc_n$N.jsonl 2> gpurun_out/r2f_check_inproc_n$N.err
cat gpurun_out/r2f_check_inproc_n$N.jsonl; tail -2 gpurun_out/r2f_check_inproc_n$N.err
timeout 300 build/bin/cfg4_multi $N 5 > gpurun_out/r2f_cfg4_c_n$N.json 2>&1; cat gpurun_out/r2f_cfg4_c_n$N.json
