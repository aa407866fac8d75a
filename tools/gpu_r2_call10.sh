#!/bin/bash
# round 2, GPU call 10: full GPU suite, then the same suite with guard bands around every plan array (bounds check; the
# pool's compute-sanitizer is closed), smoke
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/r2c10_tests.log
tail -4 gpurun_out/r2c10_tests.log
NSOL_DEBUG_GUARD=1 timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2c10_tests_guard.log
tail -6 gpurun_out/r2c10_tests_guard.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c10_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2c10_smoke.log
tail -3 gpurun_out/r2c10_smoke.log
