set -x
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --pcg-iters 8"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01e_launches.csv $B > gpurun_out/r01e_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_element_stiffness|k_assemble_gather" -s 2 -c 2 -f -o gpurun_out/r01e_k3_full $B > /dev/null 2>&1
python bench.py > gpurun_out/r01e_bench_n1.json 2> gpurun_out/r01e_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r01e_bench_reference_arm.json 2>&1
tail -c 1500 gpurun_out/r01e_bench_n1.json
