# every compute-sanitizer pass of the round over tools/sanitize_run.py; logs into gpurun_out/sanitizer/
mkdir -p gpurun_out/sanitizer
for spec in "memcheck:" "racecheck:" "synccheck:" "memcheck:pdl0" "racecheck:pdl0"; do
  tool=${spec%%:*}; arg=${spec##*:}
  name=$tool${arg:+_$arg}
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py $arg > gpurun_out/sanitizer/$name.log 2>&1
  echo "$name rc=$?"; grep -E "SUMMARY|done" gpurun_out/sanitizer/$name.log | tail -3
done
