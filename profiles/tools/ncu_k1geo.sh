B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-pcg"
TB2_PIPELINE=0 ncu --set full --clock-control none --import-source on -k regex:k_internal_force -s 6 -c 1 -f -o gpurun_out/k1geo_serial $B > /dev/null 2>&1
TB2_PIPELINE=0 TB2_K1_GEO=0 ncu --set full --clock-control none --import-source on -k regex:k_internal_force -s 6 -c 1 -f -o gpurun_out/k1plain_serial $B > /dev/null 2>&1
