for cfg in "16 1" "16 3" "16 4" "8 1" "8 3" "32 1" "32 3" "4 3"; do set -- $cfg; echo "K=$1 LB=$2"; B3_ACC_K=$1 B3_ACC_LB=$2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(' value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'accum', round(r['stage_ms']['miller_accumulate'],3), 'chain', round(r['stage_ms']['miller_chain'],3), 'e2e', round(d['e2e']['value']))"; done
