# A/B of one environment switch on the Performer step:  bash tools/ab_env.sh SA_FAVOR_FUSED_BWD   (-> gpurun_out/bench_ab_<v>.json)
VAR=${1:?variable}
for v in ${AB_RUNS:-1 0 1 0}; do env $VAR=$v python bench.py --workload performer --no-cpu-baseline --no-vendor --no-parity --no-extra --no-e2e --pf-breakdown --steps 6 --warmup 3 > gpurun_out/bench_ab_$v.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_ab_$v.json').read().strip().splitlines()[-1]); b=d['breakdown_ms']; print('$VAR=$v', round(d['ms_per_step'],2), 'loss', d['loss'], 'hbm', round(d['hbm_peak_gb'],1), {k:v['ms'] for k,v in b.items() if 'favor' in k})"; done
