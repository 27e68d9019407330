# N independent single-GPU benches at the same time (no collective): separates box-level effects (host CPU, PCIe, power) from the
# cost of meeting at the all-gather in the sharded run
N=${1:-8}
for i in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$i python bench.py --steps 5 --warmup 3 --no-next-rows --no-cpu-baseline > gpurun_out/replica_$i.json 2> gpurun_out/replica_$i.err &
done
wait
python - <<EOF
import json
for i in range($N):
    try:
        d = json.load(open("gpurun_out/replica_%d.json" % i))
        print(i, round(d["ms_per_step"], 2), round(d["value"]), round(d["e2e"]["ms_per_step"], 2))
    except Exception as e:
        print(i, "failed", e)
EOF
