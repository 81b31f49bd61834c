#!/bin/bash
# r2r: compute-sanitizer over the last changes of the round (profiles/sanitize_r2j.py)
mkdir -p gpurun_out
: > gpurun_out/r2r_sanitizer.txt
for tool in memcheck synccheck racecheck; do
  echo "== $tool" >> gpurun_out/r2r_sanitizer.txt
  timeout 110 compute-sanitizer --tool $tool python profiles/sanitize_r2j.py 2>&1 | grep -v "^$" | tail -8 >> gpurun_out/r2r_sanitizer.txt
done
cat gpurun_out/r2r_sanitizer.txt
