set -x
mkdir -p gpurun_out
nproc; lscpu | grep -E "Socket|NUMA|Model name|^CPU\(s\)" | head -8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 5 --warmup 3 --no-latency --allpairs-kf 0 --match-pairs 64 > gpurun_out/bench_n4b.json 2> gpurun_out/bench_n4b.err; tail -2 gpurun_out/bench_n4b.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n4b.json'))
print('fps', round(d['value']), 'e2e', d['e2e'], d['config'].get('numa_node_of_rank0'))
PY
