cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-sdrf --no-clocks > gpurun_out/v.json 2> gpurun_out/v.err || tail -5 gpurun_out/v.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/v.json'))
print("default lib: ms/step", round(d['ms_per_step'],4), d['config']['phase_ms_rank0'], "parity", d['config']['parity_spot_check_vs_c_oracle'])
PY
DCR_LIB_PATH=$PWD/build/libdcr_trace.so timeout 300 python profiles/paper_cta_timeline.py 2>&1 | tail -24
mv build/libdcr_trace.so build/trace.so.skip
bash profiles/jobs/r02_variants.sh
