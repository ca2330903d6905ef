#!/bin/bash
# Round-2 ncu evidence (1 GPU, under gpurun).  Numbers printed by a run under ncu are NOT bench values.
set -x
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --pcg-iters 8 --nlpcg-n 0 --no-explicit-solid --no-parity --no-shuffled"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches.csv $B > /dev/null 2> gpurun_out/r02_launches.err
ncu --set full --clock-control none --import-source on -k regex:k_internal_force_neo -s 5 -c 1 -o gpurun_out/r02_k1 $B --no-pcg > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cd_node_update -s 5 -c 1 -o gpurun_out/r02_k5 $B --no-pcg > /dev/null 2>&1
ls -la gpurun_out/r02_*
