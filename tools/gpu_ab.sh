# A/B of environment switches (join several variables of one setting with ":"): usage  gpurun -- bash tools/gpu_ab.sh "VAR=a VAR=b ..." [pytest targets]
# runs the parity tests once (default: extractor + stereo), then the short device-resident bench per setting
set -x
mkdir -p gpurun_out
SETTINGS=${1:-"X=0"}
shift
T=${@:-tests/test_gpu_extract_parity.py tests/test_gpu_stereo_parity.py}
timeout 1200 python -m pytest $T -q -x -m gpu 2>&1 | tail -8
for S in $SETTINGS; do
  env ${S//:/ } timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency --allpairs-kf 0 --match-pairs 64 > gpurun_out/b_ab.json 2> gpurun_out/b_ab.err; tail -2 gpurun_out/b_ab.err
  python -c "
import json; d=json.load(open('gpurun_out/b_ab.json')); print('AB $S fps', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 3) for k, v in d['stage_ms_per_step'].items()})"
done
