set -x
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --pcg-iters 8 --no-explicit-solid --nlpcg-n 0"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01k_launches.csv $B > gpurun_out/r01k_launches.log 2>&1
TB2_PIPELINE=0 timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_internal_force -s 6 -c 1 -f -o gpurun_out/r01k_k1_full $B --no-pcg > /dev/null 2>&1
ncu -i gpurun_out/r01k_k1_full.ncu-rep --page raw --csv > gpurun_out/r01k_k1_details.csv 2>/dev/null
ls -la gpurun_out/r01k_* | head
