#!/bin/bash
# Final single-GPU round (run under gpurun): tests, smoke, bench with all extras, the reference arm, the 1e8 point,
# ncu of the qubit kernels, compute-sanitizer on the kernels added this half of the round.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2f_smoke.log
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?" >> gpurun_out/r2f_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
timeout 300 python tests/gpu_s8.py 10000 4 > gpurun_out/r2f_s8.json 2> gpurun_out/r2f_s8.err; echo "s8 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"pauli_|csr_matvec" -c 8 -f \
    -o gpurun_out/r2prof_pauli2 python tests/gpu_qubit_driver.py solve > gpurun_out/r2prof_pauli2.out 2>&1
for TOOL in memcheck initcheck synccheck; do
    timeout 400 compute-sanitizer --tool $TOOL python -m pytest tests/test_sqd_loop.py tests/test_sigma_instances_gpu.py -q -m gpu \
        -k "carryover or (bit_array and not 70000) or (wide and 12-5-7) or (wide and spin_operator)" \
        > gpurun_out/r2f_sanitizer_$TOOL.log 2>&1; echo "$TOOL rc=$?" >> gpurun_out/r2f_sanitizer_$TOOL.log
done
tail -3 gpurun_out/r2f_tests.log; tail -2 gpurun_out/r2f_smoke.log; tail -2 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], list(d["extra"].keys()))
print(d["extra"].get("formats")); print(d["extra"].get("sqd_loop_c2"))
PY
cat gpurun_out/r2f_s8.json; for T in memcheck initcheck synccheck; do tail -4 gpurun_out/r2f_sanitizer_$T.log; done
