set -x
mkdir -p gpurun_out
nproc
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 4 --warmup 3 --no-latency --allpairs-kf 0 --match-pairs 64 > gpurun_out/bench_n8b.json 2> gpurun_out/bench_n8b.err; tail -2 gpurun_out/bench_n8b.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n8b.json'))
print('fps', round(d['value']), 'e2e', d['e2e'], d['config'].get('numa_node_of_rank0'))
PY
