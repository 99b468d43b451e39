#!/bin/bash
# Development call: full GPU suite (no -x), then A/B of library builds on the render workload.  Outputs under gpurun_out/.
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --timeout=300 ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"
  grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -25
fi
timeout 600 python tools/ab_lib.py $AB_LIBS 2>&1 | tee gpurun_out/ab_lib.log
