#!/bin/bash
# Round-2 evidence session (one B200, run under gpurun): tests, smoke, bench (both arms), ncu launch list, ncu --set full of
# every kernel family (converted to CSV on the box), host profile, micro-benchmarks.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,driver_version --format=csv,noheader > gpurun_out/gpu_r02.txt
(timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=30 > gpurun_out/gputest_r02.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_r02.log)
grep -E "^FAILED|passed|failed" gpurun_out/gputest_r02.log | tail -5
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r02.log); tail -2 gpurun_out/smoke_r02.log
(timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; echo "bench rc=$?")
(timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_r02.json 2>/dev/null; echo "reference arm rc=$?")
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python profiles/profile_step.py > gpurun_out/ncu_r02.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv3d_ring2" -c 12 -o /tmp/conv3d -f python profiles/profile_step.py >> gpurun_out/ncu_r02.log 2>&1
ncu -i /tmp/conv3d.ncu-rep --page raw --csv > gpurun_out/conv3d_r02.csv 2>/dev/null
timeout 400 ncu --profile-from-start off --set full --clock-control none -k regex:"est_attend|head_softargmin|gru_blend|gru_reset|warp_cost|premix|gn_finalize" -c 24 -o /tmp/hbm -f python profiles/profile_step.py >> gpurun_out/ncu_r02.log 2>&1
ncu -i /tmp/hbm.ncu-rep --page raw --csv > gpurun_out/hbm_kernels_r02.csv 2>/dev/null
# BASELINE configs[4]: 640x960, D=128 -- ncu roofline capture of the fused warp -> cost-volume kernel
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"warp_cost" -c 2 -o /tmp/k1cfg5 -f python profiles/profile_step.py cfg5 >> gpurun_out/ncu_r02.log 2>&1
ncu -i /tmp/k1cfg5.ncu-rep --page raw --csv > gpurun_out/warp_cost_cfg5_r02.csv 2>/dev/null
timeout 400 ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --clock-control none -k regex:"conv2d_tc" -c 130 -o /tmp/planar -f python profiles/profile_step.py >> gpurun_out/ncu_r02.log 2>&1
ncu -i /tmp/planar.ncu-rep --page raw --csv > gpurun_out/planar_r02.csv 2>/dev/null
for m in joint estm estm_ids; do timeout 120 python profiles/host_profile.py $m; done > gpurun_out/host_profile_r02.txt 2>&1
timeout 120 python profiles/bench_split.py > gpurun_out/bench_split_r02.txt 2>&1
timeout 300 python profiles/trunc_probe.py > gpurun_out/trunc_probe_r02.txt 2>&1
ls -la gpurun_out | tail -20
