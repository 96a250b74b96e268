set -x
mkdir -p gpurun_out
for cfg in "2 1" "1 1" "2 1" "1 1" "2 2"; do set -- $cfg
  FRB_OSC_MAIN_STREAMS=$1 FRB_OSC_MIN_RANGES=$2 timeout 300 python bench.py --steps 5 --warmup 3 --voices 8 --no-parity --no-extra --no-cpu-baseline > gpurun_out/r2m_8v.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2m_8v.json')); print('8 voices streams=$1 min_ranges=$2', d['ms_per_step'], d['roofline']['k1_ms_per_step'], d['roofline']['k1_family_launches_per_step'])"
done
for cfg in "2 1" "1 1" "2 1"; do set -- $cfg
  FRB_OSC_MAIN_STREAMS=$1 FRB_OSC_MIN_RANGES=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-parity --no-extra --no-cpu-baseline > gpurun_out/r2m_64v.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2m_64v.json')); print('64 voices streams=$1 min_ranges=$2', d['ms_per_step'], d['value'])"
done
