# launch list of the bench command for the final state of round 2 (per-kernel share of the serialised device time)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2D_launches.csv python bench.py --steps 1 --warmup 1 --inflight 2 --no-next-rows --no-cpu-baseline > gpurun_out/r2D_ncu_bench.log 2>&1
tail -c 200 gpurun_out/r2D_ncu_bench.log
