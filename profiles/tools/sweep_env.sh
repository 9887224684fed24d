#!/bin/bash
# usage: sweep_env.sh <out.jsonl> "<ENV=.. ENV=..>" ...   one short bench.py run per environment setting
out=$1; shift
: > "$out"
for envs in "$@"; do
  echo "== $envs" >&2
  line=$(env $envs python bench.py --steps 96 --warmup 3 --no-e2e --no-cpu-baseline --min-reps 5 --min-seconds 1 2>gpurun_out/sweep_err.log | tail -1)
  echo "{\"env\": \"$envs\", \"line\": $line}" >> "$out"
  echo "$envs -> $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"]*1e3, "us", d["roofline"]["frac"])' 2>/dev/null)"
done
