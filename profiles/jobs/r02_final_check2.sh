cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-dense > gpurun_out/r02_last_bench.json 2> gpurun_out/r02_last_bench.err || tail -5 gpurun_out/r02_last_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02_last_bench.json')); s=d['sdrf']
print('ms/step', round(d['ms_per_step'],3), 'sdrf', round(s['iters_per_s']), s['prefix_matches_cpu'], 'classical', {k:(round(v['iters_per_s']), v['prefix_matches_cpu']) for k,v in s['classical'].items()}, 'directed', round(s['directed']['iters_per_s']), s['directed']['prefix_matches_cpu'], 'numba prefix', s.get('reference_numba_on_this_gpu',{}).get('sequence_prefix_matches_ours'))"
