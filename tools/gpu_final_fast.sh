#!/bin/bash
# Refresh of the evidence after a change that touches only the fast-numerics kernel (dz_fast.cu) and host
# code: the GPU suite and smoke on the new build, and the ncu captures of the fast kernel (step-by-step
# default, and the blocked tensor-core variant).  The exact kernels' captures of tools/gpu_final_evidence.sh
# stay valid (profiles/traffic.json is keyed per kernel source file).  Output: gpurun_out/final_fast/.
O=gpurun_out/final_fast; mkdir -p $O; rm -f $O/*
sha256sum dantzig_b200/libdantzig_b200.so | cut -c1-16 > $O/lib_sha16.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
cap() { # name, kernel regex, ncu selection, args...
  name=$1; rx=$2; sel=$3; shift 3
  timeout 300 python tools/gpu_one.py "$@" > $O/one_$name.log 2>&1 && \
  timeout 600 ncu $sel --clock-control none -k regex:$rx -c 1 -o $O/ncu_$name python tools/gpu_one.py "$@" > $O/ncu_$name.log 2>&1
  ncu -i $O/ncu_$name.ncu-rep --page details > $O/ncu_${name}_details.txt 2>&1
  ncu -i $O/ncu_$name.ncu-rep --page raw --csv > $O/ncu_${name}_raw.csv 2>&1
}
cap fast_c2 dz_fast_kernel "--set full --import-source on" c2 592 0 0 0 fast
cap fast_c5 dz_fast_kernel "--set full --import-source on" c5 148 0 0 0 fast
cap fast_c5_dmma dz_fast_kernel "--set full --import-source on" c5 148 2 0 0 fast
tail -2 $O/smoke.log; cat $O/one_fast_*.log; ls $O | wc -l
