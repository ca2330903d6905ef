set -x
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --pcg-iters 10"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01d_launches.csv $B > gpurun_out/r01d_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_internal_force -s 12 -c 1 -f -o gpurun_out/r01d_k1_full $B --no-pcg > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cd_node_update -s 12 -c 1 -f -o gpurun_out/r01d_k5_full $B --no-pcg > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stiffness -s 8 -c 1 -f -o gpurun_out/r01d_k3_full $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 4 -c 1 -f -o gpurun_out/r01d_spmv_full $B > /dev/null 2>&1
ls -la gpurun_out
