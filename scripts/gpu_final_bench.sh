#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 1800 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
