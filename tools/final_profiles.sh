#!/bin/bash
# The measurements behind profiles/: tests, the driver's default bench line, launch list, ncu capture, timeline.
# Run on the GPU box through gpurun; every file lands in gpurun_out/.  tools/make_profiles.py r2 files them.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_tests.txt
python bench.py --steps 50 --warmup 5 > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err
python tools/pose_times.py > gpurun_out/final_pose_times.txt 2>&1
python tools/timeline.py > gpurun_out/final_timeline.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --legs none --sequences '' > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 12 -c 2 -f -o gpurun_out/r2_update \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --legs none --sequences '' > gpurun_out/ncu_update.log 2>&1
ls -la gpurun_out
