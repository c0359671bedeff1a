# 4-, 2- and 1-GPU runs on one 4-GPU box (weak scaling), final kernels
set -x
mkdir -p gpurun_out
for N in 4 2; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 --no-latency > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -2 gpurun_out/bench_n$N.err
done
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench_n1_4box.json 2> gpurun_out/bench_n1_4box.err
python - <<PY
import json
for n,f in ((4,'bench_n4'),(2,'bench_n2'),(1,'bench_n1_4box')):
    d=json.load(open('gpurun_out/%s.json' % f))
    print(n, 'fps', round(d['value']), 'e2e', round(d['e2e']['value']), 'stereo', round(d['stereo']['value']), 'match %.3e' % d['matching']['value'],
          'allpairs %.3e' % d.get('allpairs', {}).get('value', 0), 'track', round(d['tracking']['batch']['frames_per_s']), d['clocks']['reasons'])
PY
