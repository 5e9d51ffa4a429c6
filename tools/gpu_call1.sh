#!/bin/bash
# First GPU call: parity tests, smoke, FP32 micro-benchmark, bench, tile sweep, ncu.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== ubench"; timeout 300 ./tools/ubench_fma > gpurun_out/ubench.log 2>&1; tail -45 gpurun_out/ubench.log
echo "== sweep"; timeout 600 python tools/sweep_st.py > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
echo "== bench"; timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
echo "== ncu launches"
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
echo "== ncu full (c5 forward+backward)"
timeout 600 python bench.py --steps 4 --warmup 3 --workload c5 --no-extra --no-cpu > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:st_ -s 6 -c 4 -o gpurun_out/prof_c5 python bench.py --steps 4 --warmup 3 --workload c5 --no-extra --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
