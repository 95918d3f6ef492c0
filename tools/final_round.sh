#!/bin/bash
# end-of-round evidence: full GPU test suite, the default bench, the launch list of one timed step (ncu,
# after a plain run of the same command that exited 0), the examples
TAG=${1:-r01h}
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-220 gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_$TAG.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 336 -c 112 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
python examples/fit_hyperparameters.py 2>&1 | tail -1
python examples/fit_gpderivs.py 2>&1 | tail -2
