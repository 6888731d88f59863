#!/bin/bash
mkdir -p gpurun_out
python bench.py --taxa 200 --patterns 100000 --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?"; tail -3 gpurun_out/bench_small.err
python bench.py --impl reference --taxa 200 --patterns 100000 --steps 2 --warmup 1 --cpu-budget 20 > gpurun_out/bench_small_ref.json 2> gpurun_out/bench_small_ref.err; echo "rc=$?"; tail -3 gpurun_out/bench_small_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_small.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/bench_small_ref.json').read().strip().splitlines()[-1])
print('ours logL', d['logL'], 'ref logL', r['logL'], 'rel', abs(d['logL']-r['logL'])/abs(r['logL']))
print('ms', d['ms_per_step'], d['ms_per_step_p10'], d['ms_per_step_p50'], d['ms_per_step_p90'])
print('verify', d.get('verify'))
print('f32', {k:v for k,v in d.get('f32',{}).items() if k!='roofline'}, d.get('f32',{}).get('roofline',{}).get('scheduled_frac'))
for k in ('rbcl738','rbcl738_i_g4'):
    b=d[k]; print(k, 'dev', b['evals_per_sec_device'], 'fused', b['evals_per_sec_e2e_fused_call'], 'facade', b['evals_per_sec_e2e_cpp_facade'], 'cpu all', b['cpu_evals_per_sec_all_threads'], b['cpu_threads'], 'batched', b['batched_chains'], 'roof', b['roofline']['frac'], b['roofline']['batched_frac'])
print('mc3', d.get('mc3'))
print('ref value', r['value'], 'ours', d['value'], d['e2e']['value'])
PY
