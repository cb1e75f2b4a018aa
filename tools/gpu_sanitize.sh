#!/bin/bash
# compute-sanitizer evidence (SURVEY 4 item 5 / 5 "Race detection"): memcheck + racecheck + synccheck over the
# rasteriser, graph-step and fused-step GPU tests at C1-like shapes. Logs land in gpurun_out/<tag>/.
tag=${1:-san}
out=gpurun_out/$tag
mkdir -p $out
K="test_forward_backward_vs_oracle or test_fused_step_eager_matches_autograd or test_empty_and_invisible or test_shade_matches_reference_golden or test_fused_tail_matches_torch"
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 \
      python -m pytest tests/test_raster_gpu.py tests/test_fused_step_gpu.py tests/test_shading_gpu.py tests/test_fused_loss_gpu.py -m gpu -q -x -k "$K" > $out/$tool.log 2>&1
  echo "$tool exit $?" | tee -a $out/$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out/$tool.log | tail -4
done
