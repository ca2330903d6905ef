# the Newton leg of the PCG bench three times in one process would hide first-call effects; here: two fresh processes, PCG leg only
for k in 1 2; do
python bench.py --no-explicit-solid --nlpcg-n 0 --no-cpu-baseline --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('newton', d['pcg']['newton']['seconds'], d['pcg']['newton']['pcg_iterations'], 'pcg ms/it', d['pcg']['ms_per_iteration'])"
done
