# round-end check: what the driver runs (GPU tests, smoke, both bench arms), un-profiled
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r01k_bench_reference_arm.json 2> gpurun_out/r01k_ref.err
python bench.py > gpurun_out/r01k_bench_n1.json 2> gpurun_out/r01k_bench.err
tail -2 gpurun_out/r01k_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r01k_bench_n1.json"))
r = json.load(open("gpurun_out/r01k_bench_reference_arm.json"))
print("value %.4g e2e %.4g ref %.4g  ratio e2e/ref %.1f" % (d["value"], d["e2e"]["value"], r["value"], d["e2e"]["value"] / r["value"]))
print("pcg %.4g  assembly %s" % (d["pcg"]["value"], d["pcg"]["assembly"]["ms"]), "newton", d["pcg"]["newton"]["seconds"], "xs %.4g" % d["explicit_solid"]["value"], "j2", d["nlpcg_j2"]["element_sweeps_per_s"])
print("clocks", d["clocks"], "cpu", d["cpu_baseline"]["value"])
PY
python - <<'PY'
import json
d = json.load(open("gpurun_out/r01k_bench_n1.json"))
print("pcg cpu reference", d["pcg"].get("cpu_reference"))
PY
python profiles/tools/geom_reader_bench.py 100 2>&1 | tail -3
