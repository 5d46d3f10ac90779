#!/bin/bash
# ncu launch lists + --set full raw pages of one 3-block step at 1024 and 256 pairs, and the compute-sanitizer record
# (run under gpurun from the repo root; the .ncu-rep files stay on the box, the CSV exports come back in gpurun_out/)
set -u
for B in 1024 256; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 27 -c 26 --csv --log-file gpurun_out/launches_b$B.csv \
      python tools/profile_step.py --batch $B --steps 2 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -s 27 -c 26 -o /tmp/full_b$B -f \
      python tools/profile_step.py --batch $B --steps 2 > /dev/null 2>&1
  ncu -i /tmp/full_b$B.ncu-rep --page raw --csv > gpurun_out/full_b$B.csv 2>/dev/null
done
{
  for args in "--batch 3 --steps 1" "--batch 70 --steps 1 --variant full --show-error" "--batch 300 --steps 1"; do
    echo "compute-sanitizer --tool memcheck python tools/profile_step.py $args -> $(compute-sanitizer --tool memcheck python tools/profile_step.py $args 2>&1 | grep 'ERROR SUMMARY' | tail -1)"
  done
  for args in "--batch 3 --steps 1" "--batch 300 --steps 1"; do
    echo "compute-sanitizer --tool synccheck python tools/profile_step.py $args -> $(compute-sanitizer --tool synccheck python tools/profile_step.py $args 2>&1 | grep 'ERROR SUMMARY' | tail -1)"
  done
} > gpurun_out/sanitizer.txt 2>&1
cat gpurun_out/sanitizer.txt
wc -l gpurun_out/launches_b1024.csv gpurun_out/full_b1024.csv gpurun_out/launches_b256.csv gpurun_out/full_b256.csv
