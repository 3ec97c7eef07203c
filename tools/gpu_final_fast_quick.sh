O=gpurun_out/final_fast; mkdir -p $O
sha256sum dantzig_b200/libdantzig_b200.so | cut -c1-16 > $O/lib_sha16.txt
timeout 120 python -m pytest tests/test_fast_mode.py -q -m gpu > $O/pytest_fast_gpu.log 2>&1; tail -2 $O/pytest_fast_gpu.log
cap() { name=$1; rx=$2; sel=$3; shift 3
  timeout 100 ncu $sel --clock-control none -k regex:$rx -c 1 -o $O/ncu_$name python tools/gpu_one.py "$@" > $O/one_$name.log 2>&1
  ncu -i $O/ncu_$name.ncu-rep --page details > $O/ncu_${name}_details.txt 2>&1
  ncu -i $O/ncu_$name.ncu-rep --page raw --csv > $O/ncu_${name}_raw.csv 2>&1; }
rm -f $O/ncu_fast_c2* $O/ncu_fast_c5.* $O/ncu_fast_c5_details.txt $O/ncu_fast_c5_raw.csv
cap fast_c2 dz_fast_kernel "--set full --import-source on" c2 592 0 0 0 fast
cap fast_c5 dz_fast_kernel "--set full --import-source on" c5 148 0 0 0 fast
grep -h "^c[25] B" $O/one_fast_c2.log $O/one_fast_c5.log
