#!/usr/bin/env bash
# ncu launch list of the default bench command (per-launch durations are cold-cache and serialised: shares, not absolutes)
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/launches_bench.log 2>&1; echo "ncu rc=$?"
python profiles/summarize.py launches gpurun_out/launches.csv 2>/dev/null | head -60
