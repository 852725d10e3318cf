cd $GRAFT_REPO_ROOT
for lib in discrete-curvature-rewiring_b200/libdcr.so build/libdcr_*.so; do
echo "== $lib"
DCR_LIB_PATH=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-sdrf --no-clocks > gpurun_out/v.json 2> gpurun_out/v.err || tail -5 gpurun_out/v.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/v.json'))
print("ms/step", round(d['ms_per_step'],4), d['config']['phase_ms_rank0'], "parity", d['config']['parity_spot_check_vs_c_oracle'])
PY
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-sdrf --no-clocks"
DCR_LIB_PATH=$PWD/$lib ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__pcsamp_warps_issue_stalled_no_instructions,smsp__pcsamp_warps_issue_stalled_barrier,smsp__pcsamp_warps_issue_stalled_long_scoreboard,smsp__pcsamp_sample_buffer_full --clock-control none -k regex:'paper_(group|light)' -s 6 -c 2 --csv --log-file gpurun_out/r02_launches.csv $CMD > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_launches.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:30], r['Metric Name'][:50], r['Metric Value'])
"
done
