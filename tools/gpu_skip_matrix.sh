#!/bin/bash
# marginal machine cost of every kernel of a Davidson cycle when 8 solves share the GPU: fixed 30 cycles per
# solve (never converges), one kernel class skipped at a time (results are garbage, only the time matters)
run() { echo -n "$* : "; env "$@" python tests/gpu_solve_time.py c4 8 5 30 2>&1 | tail -1; }
run A=1
run SQD_V2_SKIP=1
run SQD_V2_SKIP=2
run SQD_V2_SKIP=4
run SQD_V2_SKIP=7
run SQD_DAV_SKIP=4
run SQD_DAV_SKIP=1
run SQD_DAV_SKIP=2
run SQD_DAV_SKIP=7
run SQD_DAV_SKIP=7 SQD_V2_SKIP=7
