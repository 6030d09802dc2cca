#!/bin/bash
# one GPU round: tests, smoke, bench, launch list (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/tests.log; tail -4 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
