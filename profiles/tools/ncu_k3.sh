B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pcg-iters 5"
TB2_K3_CHUNK=1000000 ncu --set full --clock-control none --import-source on -k regex:"k_element_stiffness|k_assemble_gather" -s 2 -c 2 -f -o gpurun_out/k3v2_1m $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_element_stiffness|k_assemble_gather" -s 144 -c 4 -f -o gpurun_out/k3v2_chunk $B > /dev/null 2>&1
python -m pytest tests -m gpu -x -q -k "two_phase or tangent" 2>&1 | tail -3
