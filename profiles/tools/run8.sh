#!/bin/bash
# the 8-GPU measurement set of one round: exchange correctness, exchange probe, C2 weak scaling (peer lag 2 / lag 0 / NCCL),
# C4 and C5 sharded.  usage: run8.sh <ngpus> <tag>
N=${1:-8}; TAG=${2:-r02}
P=29600
tr() { P=$((P+1)); python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
tr tests/multi_gpu_exchange.py > gpurun_out/${TAG}_x${N}_test.log 2>&1; tail -1 gpurun_out/${TAG}_x${N}_test.log
tr profiles/tools/exchange_probe.py 2>gpurun_out/${TAG}_x${N}_probe.err | grep -E "us / step|status|FAILED" > gpurun_out/${TAG}_exchange_probe_${N}gpu.txt; cat gpurun_out/${TAG}_exchange_probe_${N}gpu.txt
B="--gpus $N --min-seconds 1 --min-reps 5"
tr bench.py $B 2>gpurun_out/${TAG}_b1.err | tail -1 > gpurun_out/${TAG}_bench_c2_${N}gpu.json
tr bench.py $B --no-e2e --lag 0 2>gpurun_out/${TAG}_b2.err | tail -1 > gpurun_out/${TAG}_bench_c2_${N}gpu_lag0.json
tr bench.py $B --no-e2e --exchange nccl 2>gpurun_out/${TAG}_b3.err | tail -1 > gpurun_out/${TAG}_bench_c2_${N}gpu_nccl.json
tr bench.py $B --no-e2e --config c4 --sharded 2>gpurun_out/${TAG}_b4.err | tail -1 > gpurun_out/${TAG}_bench_c4_${N}gpu_sharded.json
tr bench.py $B --no-e2e --config c5 --sharded 2>gpurun_out/${TAG}_b5.err | tail -1 > gpurun_out/${TAG}_bench_c5_${N}gpu_sharded.json
for f in gpurun_out/${TAG}_bench_*_${N}gpu*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], "%.2f us/step" % (d["ms_per_step"]*1e3), "%.3g p-steps/s" % d["value"], d["config"].get("exchange_selfcheck"))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
