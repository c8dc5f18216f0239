#!/bin/bash
# bench (small + default), ncu launch list, one ncu --set full capture of the chain kernel
mkdir -p gpurun_out
U=${U:-8000000}
python bench.py --units $U --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
tail -c 3000 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --units 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fp_chain -s 1 -c 1 -o gpurun_out/prof_chain \
    python bench.py --units 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
