set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_oscbank.py tests/test_full_size.py tests/test_stream.py tests/test_edge_cases.py -x -q -m gpu 2>&1 | tail -2
for cfg in "2 4" "1 4" "1 1" "2 8" "2 4"; do set -- $cfg
  FRB_OSC_MAIN_STREAMS=$1 FRB_OSC_MIN_RANGES=$2 timeout 300 python bench.py --steps 5 --warmup 3 --voices 8 --no-parity --no-extra --no-cpu-baseline > gpurun_out/r2m_8v.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2m_8v.json')); print('8 voices streams=$1 min_ranges=$2', d['ms_per_step'], d['roofline']['k1_ms_per_step'], d['roofline']['k1_family_launches_per_step'])"
done
for cfg in "2 4" "1 1" "2 8" "2 4" "1 1"; do set -- $cfg
  FRB_OSC_MAIN_STREAMS=$1 FRB_OSC_MIN_RANGES=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2m_64v.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2m_64v.json')); print('64 voices streams=$1 min_ranges=$2', d['ms_per_step'], d['value'], d['parity']['max_err_of_full_scale'])"
done
