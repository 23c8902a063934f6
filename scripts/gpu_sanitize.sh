#!/bin/bash
# compute-sanitizer passes over small GPU configurations (memcheck / racecheck / synccheck "on the
# small configs").  Not run in round 1 (the GPU budget was spent on measurements); first thing to run in round 2:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize.sh'
# Output: gpurun_out/sanitize_<tool>.log (the tail of each log is printed).
set -u
mkdir -p gpurun_out
SEL='cuda_full_protocol_bitwise and (funnel or logit) or tensor_gradient_within_tolerance or synthetic_rows_on_device or cuda_resume_is_bitwise'
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 99 --launch-timeout 120 \
    python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: exit $?"; tail -n 6 gpurun_out/sanitize_$tool.log
done
