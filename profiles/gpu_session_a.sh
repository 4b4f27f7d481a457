#!/bin/bash
# One GPU session: gpu tests, smoke, bench, ncu launch list + --set full capture of the HBM-bound kernels (run under gpurun).
mkdir -p gpurun_out
TAG=${1:-a}
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
(timeout 1200 python -m pytest tests -m gpu -q -s --maxfail=30 > gpurun_out/gputest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_$TAG.log)
tail -5 gpurun_out/gputest_$TAG.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log)
tail -2 gpurun_out/smoke_$TAG.log
(timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?")
tail -c 600 gpurun_out/bench_$TAG.err
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_$TAG.csv python profiles/profile_step.py > gpurun_out/ncu_$TAG.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"est_attend|head_softargmin|gru_blend|gru_reset|warp_cost|premix" -c 10 -o gpurun_out/hbm_kernels_r02_$TAG -f python profiles/profile_step.py >> gpurun_out/ncu_$TAG.log 2>&1
ls -la gpurun_out | tail -12
