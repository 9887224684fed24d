#!/bin/bash
# usage: measure_variants.sh <tag> <lib1.so> [<lib2.so> ...] [-- "ENV=.. ENV=.." ...]
# For every library (T2D_B200_LIB) x environment setting: one short bench.py run (CUDA-graph timing) and one ncu
# counter pass of a single tick launch (instructions executed, issue utilisation, registers).  Results: gpurun_out/<tag>.txt
tag=$1; shift
libs=(); envs=("T2D_X=0")
while [ $# -gt 0 ]; do
  if [ "$1" == "--" ]; then shift; envs=("$@"); break; fi
  libs+=("$1"); shift
done
out=gpurun_out/$tag.txt
: > "$out"
cfg=${T2D_SWEEP_CONFIG:-c2}
for lib in "${libs[@]}"; do
  for e in "${envs[@]}"; do
    line=$(env T2D_B200_LIB=$PWD/$lib $e python bench.py --config $cfg --steps 96 --warmup 3 --no-e2e --no-cpu-baseline --min-reps 5 --min-seconds 1 2>>gpurun_out/$tag.err | tail -1)
    us=$(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("%.2f us frac %.4f" % (d["ms_per_step"]*1e3, d["roofline"]["frac"]))' 2>/dev/null)
    env T2D_B200_LIB=$PWD/$lib $e ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,gpu__time_duration.sum,smsp__warps_eligible.avg.per_cycle_active \
       --clock-control none -k regex:t2d_step_kernel -s 4 -c 1 --csv python bench.py --config $cfg --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --min-reps 1 --max-reps 1 --replicas 4 2>>gpurun_out/$tag.err \
       | grep -E "inst_executed|issue_active|registers|time_duration|warps_eligible" | awk -F'","' '{gsub(/"/,"",$NF); printf "%s=%s ", $(NF-2), $NF}' > gpurun_out/.m.txt
    echo "$lib [$e] $us | $(cat gpurun_out/.m.txt)" | tee -a "$out"
  done
done
