# BASELINE configs 3 and 5 at 8 GPUs: EuRoC-shape extraction sharded over 8 ranks, all-pairs over 4096 keyframes x 1000 descriptors
set -x
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --workload euroc --allpairs-kf 4096 --steps 4 --warmup 3 --no-latency > gpurun_out/bench_n8_euroc_full.json 2> gpurun_out/bench_n8_euroc_full.err; tail -3 gpurun_out/bench_n8_euroc_full.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n8_euroc_full.json'))
print('fps', d['value'], 'e2e', d['e2e']['value'], 'allpairs', d['allpairs'], 'match', d['matching']['value'], d['clocks'])
PY
