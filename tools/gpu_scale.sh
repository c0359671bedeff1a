# usage: gpurun --gpus N -- bash tools/gpu_scale.sh N   (weak-scaling check of bench.py at N and 1 GPUs)
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -5 gpurun_out/bench_n$N.err
python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<PY
import json
for n in ($N, 1):
    d=json.load(open('gpurun_out/bench_n%d.json' % n))
    print(n, 'fps', d['value'], 'e2e', d['e2e']['value'], 'h2d', d['e2e'].get('h2d_GBps_copy_engine_alone'), 'stereo', d['stereo']['value'], 'match', d['matching']['value'],
          'allpairs', d.get('allpairs', {}).get('value'), 'track', d['tracking']['batch']['frames_per_s'], 'clocks', d['clocks'])
PY
