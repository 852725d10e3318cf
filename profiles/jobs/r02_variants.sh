# time every build/libdcr_*.so variant with the default bench (arxiv shape, spot parity check inside)
cd $GRAFT_REPO_ROOT
for lib in build/libdcr_*.so; do
  DCR_LIB_PATH=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-sdrf --no-clocks > gpurun_out/v.json 2> gpurun_out/v.err || { echo "$lib FAILED"; tail -3 gpurun_out/v.err; continue; }
  python - "$lib" <<'PY'
import json,sys
d=json.load(open('gpurun_out/v.json'))
print(sys.argv[1], "ms/step", round(d['ms_per_step'],4), d['config']['phase_ms_rank0'], "parity", d['config']['parity_spot_check_vs_c_oracle'])
PY
done
