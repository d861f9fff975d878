#!/bin/bash
# ncu launch list of training steps (profiles/train_step_probe.py): per-launch device time, cold-cache and serialised
set -e
python profiles/train_step_probe.py 4 > gpurun_out/train_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_train.csv python profiles/train_step_probe.py 4 > gpurun_out/train_ncu.log 2>&1
