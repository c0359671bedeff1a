# 8-GPU weak-scaling point (one box): N=8 then N=1 on the same box, short runs, no CPU baseline
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --no-latency > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -3 gpurun_out/bench_n8.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench_n1_8box.json 2> gpurun_out/bench_n1_8box.err
python - <<PY
import json
for n,f in ((8,'bench_n8'), (1,'bench_n1_8box')):
    d=json.load(open('gpurun_out/%s.json' % f))
    print(n, 'fps', d['value'], 'e2e', d['e2e']['value'], 'stereo', d['stereo']['value'], 'match', d['matching']['value'],
          'allpairs', d.get('allpairs', {}).get('value'), 'track', d['tracking']['batch']['frames_per_s'], 'clocks', d['clocks'])
PY
