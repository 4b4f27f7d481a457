#!/bin/bash
# Quick GPU session: gpu tests + bench (no ncu).  Run under gpurun:  bash profiles/gpu_session_b.sh TAG
mkdir -p gpurun_out
TAG=${1:-q}
(timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=30 > gpurun_out/gputest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest_$TAG.log)
grep -E "^FAILED|passed|failed" gpurun_out/gputest_$TAG.log | tail -12
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log)
tail -2 gpurun_out/smoke_$TAG.log
(timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?")
tail -c 800 gpurun_out/bench_$TAG.err
