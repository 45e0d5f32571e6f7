#!/bin/bash
mkdir -p gpurun_out
T="tests/test_golden_gpu.py::test_reference_style_loop_matches_fused_step"
for i in 1 2 3; do timeout 300 python -m pytest "$T" -x -q 2>&1 | grep -E "passed|failed|AssertionError:" ; done
echo "--- two-pass infonce"; SBR_INFONCE_TWO_PASS=1 timeout 300 python -m pytest "$T" -x -q 2>&1 | grep -E "passed|failed|AssertionError:"
echo "--- scalar norm"; SBR_NORM_SCALAR=1 timeout 300 python -m pytest "$T" -x -q 2>&1 | grep -E "passed|failed|AssertionError:"
echo "--- both"; SBR_INFONCE_TWO_PASS=1 SBR_NORM_SCALAR=1 timeout 300 python -m pytest "$T" -x -q 2>&1 | grep -E "passed|failed|AssertionError:"
