#!/bin/bash
# Round-2 GPU call 14: ncu launch list of the final default path (shares of the step), timing of an upsampled + noisy frame.
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final_8k-d1.csv \
    python tools/profile_run.py 8k-d1 3 f32 > gpurun_out/ncu_list_final.log 2>&1
tail -2 gpurun_out/ncu_list_final.log
timeout 400 python tools/measure_post_stages.py 2>&1 | tail -6
