#!/bin/bash
# Dev tool (CPU): ThreadSanitizer sweep of the 32-lane host simulation (tests/_build/lanes32_tsan, built by the tests) over
# profiles x block sizes x presets x content kinds; prints every run that reports a race, hangs or times out.
exe=/root/repo/tests/_build/lanes32_tsan
n=0; bad=0
for prof in 1 0 3 2; do
for blk in "4 4" "5 4" "6 5" "6 6" "8 6" "8 8" "10 8" "12 12"; do
for q in 10 60 98 100; do
for kind in 0 1 2; do
  # HDR content only with HDR profiles, LDR content with LDR profiles
  if [ $kind = 2 ] && [ $prof -lt 2 ]; then continue; fi
  if [ $kind != 2 ] && [ $prof -ge 2 ]; then continue; fi
  set -- $blk
  w=$(( $1 * 2 + 1 )); h=$(( $2 * 2 ))
  out=$(TSAN_OPTIONS="halt_on_error=1 exitcode=66" timeout 300 $exe $prof $1 $2 $q $w $h $((n+7)) $kind 2>&1)
  rc=$?
  n=$((n+1))
  if [ $rc != 0 ] || echo "$out" | grep -q ThreadSanitizer; then bad=$((bad+1)); echo "FAIL prof=$prof blk=$blk q=$q kind=$kind rc=$rc"; echo "$out" | grep -E "WARNING|#0|simt_emul" | head -5; fi
done; done; done; done
echo "runs $n bad $bad"
