#!/bin/bash
# Round profile on the GPU box (run through gpurun): bench line, ncu launch list of the bench command, one ncu --set full
# capture of the first kernels of a forward (one of each kind incl. the four encoder GEMMs).  Outputs go to gpurun_out/.
#   bash tools/profile_round.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err || exit 1
python bench.py --impl reference --steps 1 --warmup 0 --cpu-budget 10 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
# launch list of the same command (durations are cold-cache and serialised: shares, not absolutes, are comparable)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
# full capture: the first 16 launches of one forward of a 61440-row batch
ncu --set full --clock-control none --import-source on -c 16 -f -o $OUT/${TAG}_full python tools/prof_run.py 61440 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -8
