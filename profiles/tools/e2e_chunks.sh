run() { env "$@" python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-pcg --no-explicit-solid | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', 'e2e %.4g'%d['e2e']['value'], 'ms/step %.4f'%d['e2e']['ms_per_step'])"; }
run TB2_HOST_CHUNKS=4
run TB2_HOST_CHUNKS=6
run TB2_HOST_CHUNKS=8
run TB2_HOST_CHUNKS=12
run TB2_HOST_CHUNKS=16
run TB2_HOST_CHUNKS=24
run TB2_HOST_CHUNKS=32
