cd $GRAFT_REPO_ROOT
python profiles/range_launches.py && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_range8_launches.csv python profiles/range_launches.py > /dev/null 2>&1
python profiles/summarize.py launches gpurun_out/r02_range8_launches.csv gpurun_out/r02_range8_launches_summary.csv | head -20
