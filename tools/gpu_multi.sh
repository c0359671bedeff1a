set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
for k in ('value','n_gpus','ms_per_step','e2e','stage_ms_per_step','allpairs'):
    print(k, d.get(k))
print(d['matching']['value'])
PY
python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
for k in ('value','n_gpus','ms_per_step','e2e','stage_ms_per_step','allpairs','cpu_baseline'):
    print(k, d.get(k))
print(d['matching']['value'])
PY
