for S in 16 24 32; do
sed -i "s/^constexpr int kDescSlots = [0-9]*;/constexpr int kDescSlots = $S;/" orb_slam2_detailed_comments_b200/csrc/orb_extract.cu
make -C orb_slam2_detailed_comments_b200/csrc > /dev/null 2>&1
for L in 2 1; do
ORB_B200_LANES=$L python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --allpairs-kf 0 --match-pairs 64 --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('SLOTS=$S LANES=$L', round(d['value']), round(d['stereo']['value']), round(d['stage_timing']['ms_per_step_serialised'],2), {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})"
done
done
