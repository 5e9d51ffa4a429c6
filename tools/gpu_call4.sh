#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
echo "== sweep"; SWEEP_FWD=0,1,2,3 timeout 600 python tools/sweep_st.py > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
echo "== bench"; timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== ncu launch list (c2)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'st_|bb_' -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
