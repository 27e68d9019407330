ncu --set full --clock-control none --import-source on -k 'regex:k_miller_accum' -c 1 -o gpurun_out/r2y_accum32768 -f python profiles/run_one.py 1 32768 > gpurun_out/r2y_accum.log 2>&1
tail -2 gpurun_out/r2y_accum.log
