python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "TB2_K1_GEO=0" "TB2_K1_GEO_MINBLOCKS=2" "TB2_K1_GEO_MINBLOCKS=3" "TB2_K1_GEO_MINBLOCKS=4"; do
env $v python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-pcg | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'k1 launch ms %.4f'%d['roofline']['avg_launch_ms'], 'e2e %.4g'%d['e2e']['value'])"
done
for v in "TB2_K1_GEO=0" "TB2_K1_GEO_MINBLOCKS=2" "TB2_K1_GEO_MINBLOCKS=3" "TB2_K1_GEO_MINBLOCKS=4"; do
env TB2_PIPELINE=0 $v python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-pcg | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('serial $v', 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'k1 launch ms %.4f'%d['roofline']['avg_launch_ms'])"
done
