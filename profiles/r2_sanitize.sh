compute-sanitizer --tool memcheck --error-exitcode 1 python profiles/memcheck_run.py > gpurun_out/r2_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_memcheck.txt
tail -4 gpurun_out/r2_memcheck.txt
compute-sanitizer --tool racecheck --error-exitcode 1 python profiles/memcheck_run.py > gpurun_out/r2_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_racecheck.txt
tail -4 gpurun_out/r2_racecheck.txt
