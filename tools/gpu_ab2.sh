for FW in 4 6 11; do
for L in 2 1; do
ORB_B200_FAST_WARPS=$FW ORB_B200_LANES=$L python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --allpairs-kf 0 --match-pairs 64 --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('FASTW=$FW LANES=$L', round(d['value']), round(d['stereo']['value']), round(d['stage_timing']['ms_per_step_serialised'],2), {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})"
done
done
