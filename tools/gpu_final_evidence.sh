#!/bin/bash
# Round-2 evidence on the FINAL build, part A: the GPU test suite, smoke, the ncu launch list of the
# default bench command and ncu captures of every kernel (smaller launches of the same kernels; the
# long-running batched kernels with a reduced section list: a full set replays the kernel ~40 times).
# Everything lands in gpurun_out/final/.  Then, here: python tools/make_traffic.py (profiles/traffic.json,
# keyed by the library hash), and part B: tools/gpu_final_bench.sh (the bench lines, which then carry
# roofline.traffic of this very library).
O=gpurun_out/final; mkdir -p $O; rm -f $O/*
sha256sum dantzig_b200/libdantzig_b200.so | cut -c1-16 > $O/lib_sha16.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 1 > $O/bench_c5_for_launches.json 2> $O/bench_c5_for_launches.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c5.csv \
    python bench.py --steps 2 --warmup 1 > $O/ncu_launches.log 2>&1
SECS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section SchedulerStats --section WarpStateStats --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
cap() { # name, kernel regex, ncu selection, args...
  name=$1; rx=$2; sel=$3; shift 3
  timeout 300 python tools/gpu_one.py "$@" > $O/one_$name.log 2>&1 && \
  timeout 600 ncu $sel --clock-control none -k regex:$rx -c 1 -o $O/ncu_$name python tools/gpu_one.py "$@" > $O/ncu_$name.log 2>&1
  ncu -i $O/ncu_$name.ncu-rep --page details > $O/ncu_${name}_details.txt 2>&1
  ncu -i $O/ncu_$name.ncu-rep --page raw --csv > $O/ncu_${name}_raw.csv 2>&1
}
cap warp_c5 dz_batch_kernel "$SECS" c5 592 -1
cap warp_c2 dz_batch_kernel "$SECS" c2 2048 -1
cap core_c5 dz_core_kernel "$SECS" c5 296 0 4 2
cap fast_c2 dz_fast_kernel "--set full --import-source on" c2 592 0 0 0 fast
cap fast_c5 dz_fast_kernel "--set full --import-source on" c5 148 0 0 0 fast
cap fast_c5_dmma dz_fast_kernel "--set full --import-source on" c5 148 2 0 0 fast
cap grid_c4 dz_grid_kernel "--set full --import-source on" c4 60
cap grid_c3 dz_grid_kernel "--set full --import-source on" c3 60
tail -2 $O/smoke.log 2>/dev/null | tail -4; ls $O | wc -l
