#!/bin/bash
# Round-2 ncu evidence, part b (run under gpurun, ONE GPU): `--set full` captures of the kernels outside the
# fermion sigma build -- Pauli projection, CSR matvec, key sort, configuration recovery -- and of the wide sigma
# kernel at a 8.1e7-determinant subspace.  Numbers printed by these runs are never bench values.
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"pauli_|csr_matvec|bitonic_|key_table_build" -c 14 -f \
    -o gpurun_out/r2prof_qubit python tests/gpu_qubit_driver.py solve > gpurun_out/r2prof_qubit.out 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"recover_|merge_" -c 16 -f \
    -o gpurun_out/r2prof_recover python tests/gpu_recovery_driver.py exact > gpurun_out/r2prof_recover.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sigma_wide" -c 1 -f \
    -o gpurun_out/r2prof_wide python tests/gpu_s8.py 9000 2 > gpurun_out/r2prof_wide.out 2>&1
ls -la gpurun_out | grep -E "r2prof_(qubit|recover|wide)"
tail -2 gpurun_out/r2prof_wide.out
