#!/bin/bash
# Round-2 first GPU call: box probe, A/B of the emulator-validated switches, phase profile,
# launch list + one full ncu capture of the shipped warp kernel.
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
{ nproc; nvidia-smi -L; ls -d /root/reference baseline/_ref 2>&1; which cargo rustc go node 2>&1; free -g | head -2; } > $O/probe.txt 2>&1
AB_FAST=1 timeout 900 python tools/gpu_ab.py build/ab/base.so build/ab/tiled.so build/ab/bsub.so build/ab/tb.so build/ab/pernr.so build/ab/tbp.so build/ab/noinl.so build/ab/opq.so build/ab/base.so > $O/ab.log 2>&1
DZ_LIB=$PWD/build/ab/prof.so timeout 300 python tools/gpu_prof.py 4096 auto > $O/prof.log 2>&1
timeout 300 python tools/gpu_one.py 4096 > $O/one.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dz_batch_kernel -c 1 -o $O/r02_shipped_warp_c2 python tools/gpu_one.py 4096 > $O/ncu_full.log 2>&1
ncu -i $O/r02_shipped_warp_c2.ncu-rep --page details > $O/r02_shipped_warp_c2_details.txt 2>&1
tail -3 $O/ab.log $O/prof.log $O/one.log
