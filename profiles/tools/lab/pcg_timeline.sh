#!/bin/bash
# Lab: device-side timeline of the distributed PCG iteration.  Builds a second copy of the library with -DTB2_PCG_TIMELINE
# (event marks inside iterations 16..23 of pcg_distributed) into profiles/tools/lab/_lib/ (git-ignored; built here, it travels with the snapshot), and runs the PCG leg of bench.py on N GPUs with it.
#   usage (GPU box): profiles/tools/lab/pcg_timeline.sh 2
set -e
N=${1:-2}
ROOT=$(cd "$(dirname "$0")/../../.." && pwd)
[ -f $ROOT/profiles/tools/lab/_lib/libtahoe_b200.so ] || make -s -j8 -C $ROOT/tahoe_b200/csrc OUT=$ROOT/profiles/tools/lab/_lib OBJ=$ROOT/build/lab_tl/obj \
  NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DTB2_PCG_TIMELINE" > /dev/null
cd $ROOT
TB2_LAB_LIB=$ROOT/profiles/tools/lab/_lib/libtahoe_b200.so python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
  profiles/tools/lab/pcg_timeline.py
