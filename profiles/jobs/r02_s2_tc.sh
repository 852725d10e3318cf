cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bfc.py tests/test_gpu_dropin.py tests/test_gpu_ref_kernels.py -x -q -k "tensor or dense or squirrel" 2>&1 | tail -3
timeout 600 python bench.py --workload squirrel-dense --steps 10 --warmup 3 2>gpurun_out/tc.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps(d)[:1800])
"
