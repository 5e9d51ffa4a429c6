#!/bin/bash
# Round-end style run: all GPU tests, smoke, both bench arms, ncu launch list of the bench command.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-1500; tail -3 gpurun_out/bench_ref.err
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-extra --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"; grep -c st_ gpurun_out/launches.csv
fi
