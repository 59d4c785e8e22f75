# Round-end refresh on the GPU box: full GPU test suite, smoke(), bench line, ncu launch list, engine/MC timings.
# Usage: gpurun --timeout 1000 -- bash tools/round_end_refresh.sh ; then copy gpurun_out/* into profiles/.
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r1c.log 2>&1; tail -3 gpurun_out/pytest_gpu_r1c.log
timeout -k 5 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -k 5 240 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -c 600 gpurun_out/bench_r1c.json
timeout -k 5 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; wc -l gpurun_out/launches_r1c.csv
timeout -k 5 120 python tools/bench_engine.py > gpurun_out/engine_r1c.json 2>&1; tail -5 gpurun_out/engine_r1c.json
