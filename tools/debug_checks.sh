#!/bin/bash
# on the GPU box: the whole GPU test suite against a build with device-side index checks (-DPT_DEBUG_CHECKS=1: output slots,
# queue slots, pixel indices, traversal stack, deferred lists, compaction positions; a violation prints and traps).
# compute-sanitizer is closed on the B200 pool; this is the substitute.  Log -> gpurun_out/r02_debug_checks.txt
cd "$(dirname "$0")/.."
[ -f build/libpt_dbg.so ] || tools/build_variant.sh dbg -DPT_DEBUG_CHECKS=1 > /dev/null
log=gpurun_out/r02_debug_checks.txt
{
  echo "# PT_B200_LIB=build/libpt_dbg.so (nvcc ... -DPT_DEBUG_CHECKS=1): pytest -m gpu + tools/sanitize_driver.py, both bounce kernels"
  PT_B200_LIB=build/libpt_dbg.so python -m pytest tests -m gpu -q 2>&1 | tail -5
  echo "# PT_B200_FUSED=1 (the fused kernel at every depth)"
  PT_B200_FUSED=1 PT_B200_LIB=build/libpt_dbg.so python -m pytest tests -m gpu -q 2>&1 | tail -5
  echo "# tools/sanitize_driver.py"
  PT_B200_LIB=build/libpt_dbg.so python tools/sanitize_driver.py 2>&1 | tail -30
  echo "# PT_CHECK failures printed by the device: $(grep -c 'PT_CHECK failed' $log 2>/dev/null || echo 0)"
} > $log 2>&1
tail -12 $log
