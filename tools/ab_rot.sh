# A/B of SA_LOCAL_ROT_FUSED (the rotary transpose inside the local-attention backward epilogues vs its own pass)
for v in 1 0 1 0; do SA_LOCAL_ROT_FUSED=$v python bench.py --workload performer --no-cpu-baseline --no-vendor --no-parity --no-extra --no-e2e --pf-breakdown --steps 6 --warmup 3 > gpurun_out/bench_rot$v.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_rot$v.json').read().strip().splitlines()[-1]); b=d['breakdown_ms']; print('fused=$v', round(d['ms_per_step'],2), {k:v['ms'] for k,v in b.items() if 'local' in k or 'rotary' in k})"; done
