#!/bin/bash
# Race / memory checks for the sm_100a kernels (run on a GPU box).  compute-sanitizer does not
# see cross-device races: those are covered by tests/test_gpu_engine.py::test_two_gpu_* (replicas must stay
# bit-identical) and by the step-stamped flag protocol (flags are never reset, only compared).
set -e
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x \
    -k "gram or full_rank or ps_update or qsgd or entrywise or tcgen05" -p no:cacheprovider
done
