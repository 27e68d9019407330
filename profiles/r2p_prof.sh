# launch list of the bench command (per-kernel share of the serialised device time) + full captures of the dominant kernels at the C4 shape
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 1 --warmup 1 --inflight 2 --no-next-rows --no-cpu-baseline > gpurun_out/r2p_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r2p_ncu_bench.log
ncu --set full --clock-control none --import-source on -k 'regex:k_hash_to_g2|k_miller_accum|k_g1_aggregate_idx|k_miller_lines_q|k_g1_mul_u64_pp_d' --launch-skip 1 -c 6 -o gpurun_out/r2p_c4 -f python profiles/run_latency.py 1 > gpurun_out/r2p_c4.log 2>&1
tail -3 gpurun_out/r2p_c4.log
