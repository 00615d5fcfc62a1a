for t in 3552 5328 7104 10656 14208; do
  SOROBN_B200_TARGET_CTAS=$t python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('target',$t,d['ms_per_step'])"
done
for l in 1024 2048 8192; do
  SOROBN_B200_LIFT_MAX=$l python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('lift',$l,d['ms_per_step'])"
done
SOROBN_B200_PDL=1 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('pdl',d['ms_per_step'])"
