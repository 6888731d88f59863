#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_ref.err
( time python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1])
print('ours logL', d['logL'], 'ref logL', r['logL'], 'rel', abs(d['logL']-r['logL'])/abs(r['logL']))
print('ms', d['ms_per_step'], d['ms_per_step_p10'], d['ms_per_step_p50'], d['ms_per_step_p90'], 'clocks', d['clocks'])
print('roofline', d['roofline'])
print('verify', d.get('verify'))
print('f32', d.get('f32'))
for k in ('rbcl738','rbcl738_i_g4'):
    b=d[k]; print(k, 'dev', b['evals_per_sec_device'], 'fused', b['evals_per_sec_e2e_fused_call'], 'facade', b['evals_per_sec_e2e_cpp_facade'], 'cpu all', b['cpu_evals_per_sec_all_threads'], b['cpu_threads'], 'batched', b['batched_chains'], 'roof', b['roofline']['frac'], b['roofline']['batched_frac'])
print('mc3', d.get('mc3'))
print('ref value', r['value'], r['steps'], r['ms_per_step'], 'ours', d['value'], d['e2e']['value'], 'ratio', d['e2e']['value']/r['value'])
PY
