python -m pytest tests -m gpu -x -q -k "two_phase or tangent or newton" 2>&1 | tail -3
for v in "" "TB2_K3_CHUNK=127872" "TB2_K3_CHUNK=504384" "TB2_K3_CHUNK=1000000"; do
env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --pcg-iters 20 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['pcg']['assembly'])"
done
