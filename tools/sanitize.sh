#!/bin/bash
# on the GPU box: compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over every kernel family of the library,
# small sizes (tools/sanitize_driver.py).  Logs -> gpurun_out/r02_sanitizer_<tool>[_fused].txt (copied to profiles/).
# usage: tools/sanitize.sh [tool ...]
cd "$(dirname "$0")/.."
tools=${@:-memcheck racecheck synccheck initcheck}
rc=0
for t in $tools; do
  for fused in 0 1; do
    tag=$t; [ $fused = 1 ] && tag=${t}_fused
    # the second pass runs the fused kernel at every depth (the A/B knob), so that both bounce kernels are covered
    [ $fused = 1 ] && [ $t != memcheck ] && [ $t != racecheck ] && continue
    log=gpurun_out/r02_sanitizer_$tag.txt
    PT_B200_FUSED=$fused timeout 1500 compute-sanitizer --tool $t --error-exitcode 9 --print-limit 20 \
      python tools/sanitize_driver.py > $log 2>&1
    e=$?
    echo "exit code $e" >> $log
    echo "$tag: exit $e, $(grep -c 'ERROR SUMMARY' $log) summary line(s): $(grep 'ERROR SUMMARY' $log | tail -1)"
    [ $e != 0 ] && rc=1
  done
done
exit $rc
