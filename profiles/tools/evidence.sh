#!/bin/bash
# The evidence set of a round on ONE B200 (tag = rNN_vK): bench lines of the four configurations, the launch list of the bench
# command, one `ncu --set full` capture of the tick (raw metrics, per-line hot spots, SASS-level stalls, DRAM traffic), the
# phase timeline, the batch-size sweep, the auxiliary kernels.   usage: evidence.sh <tag>
TAG=${1:-r02_v6}
O=gpurun_out
python bench.py > $O/${TAG}_bench_c2.json 2>$O/${TAG}_bench_c2.err
for c in c3 c4 c5; do python bench.py --config $c --no-e2e --no-cpu-baseline --steps 24 --min-seconds 1 2>/dev/null | tail -1 > $O/${TAG}_bench_$c.json; done
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/${TAG}_bench_c2_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c2.csv python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --min-reps 1 --max-reps 1 --min-seconds 0 --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:t2d_step_kernel -s 4 -c 1 -f -o $O/${TAG}_step python bench.py --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --min-reps 1 --max-reps 1 --replicas 4 > /dev/null 2>&1
python profiles/tools/size_sweep.py > $O/${TAG}_size_sweep.txt 2>&1
python profiles/tools/aux_kernels.py 2>&1 | tail -9 > $O/${TAG}_aux_kernels.txt
python profiles/tools/ablate.py > $O/${TAG}_ablation.txt 2>&1
if [ -f build/variants/dbg.so ]; then T2D_B200_LIB=$PWD/build/variants/dbg.so python profiles/phase_clocks.py > $O/${TAG}_phase_clocks_c2.txt 2>&1; fi
ls -la $O | grep ${TAG}
