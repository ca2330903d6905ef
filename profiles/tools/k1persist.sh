python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { env "$@" python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-pcg | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'k1 launch ms %.4f'%d['roofline']['avg_launch_ms'])"; }
run TB2_K1_PERSIST=0
run TB2_K1_PERSIST=1
run TB2_K1_PERSIST=2
run TB2_K1_PERSIST=1 TB2_K1_MINBLOCKS=2
run TB2_K1_PERSIST=1 TB2_K1_MINBLOCKS=2 TB2_K5_THREADS=128
run TB2_K1_PERSIST=1 TB2_K1_MINBLOCKS=4
run TB2_PIPELINE=0 TB2_K1_PERSIST=0
run TB2_PIPELINE=0 TB2_K1_PERSIST=1
run TB2_PIPELINE=0 TB2_K1_PERSIST=1 TB2_K1_MINBLOCKS=2
run TB2_PIPELINE=0 TB2_K1_PERSIST=1 TB2_K1_MINBLOCKS=4
