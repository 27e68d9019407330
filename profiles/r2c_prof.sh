set -x
ncu --set full --clock-control none --import-source on -k 'regex:k_hash_to_g2|k_miller_lines|k_miller_accum|k_g1_aggregate|k_g2_subgroup|k_msm_bucket|k_g1_mul_u64_pp' -c 9 -o gpurun_out/r2c_big -f python profiles/run_one.py 1 32768 > gpurun_out/r2c_big.log 2>&1
tail -3 gpurun_out/r2c_big.log
