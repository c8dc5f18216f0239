#!/bin/bash
# round-end style measurement: parity tests, reference arm, default bench (full BASELINE size), other workloads, ncu summaries
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 2500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
python bench.py --workload pe150_full --units 20000000 --no-cpu-baseline --fastq-units 0 > gpurun_out/bench_pe150_full.json 2> gpurun_out/bench_pe150_full.err; python -c "
import json;d=json.load(open('gpurun_out/bench_pe150_full.json'));print('pe150_full', d['value']/1e6, d['e2e']['value']/1e6, d['checks'])"
python bench.py --workload se150_cut_right_polyg --no-cpu-baseline --fastq-units 0 > gpurun_out/bench_se150.json 2> gpurun_out/bench_se150.err; python -c "
import json;d=json.load(open('gpurun_out/bench_se150.json'));print('se150', d['value']/1e6, d['e2e']['value']/1e6, d['checks'])"
python bench.py --profile 0 --units 20000000 --no-cpu-baseline --fastq-units 0 > gpurun_out/bench_refstyle.json 2> gpurun_out/bench_refstyle.err; python -c "
import json;d=json.load(open('gpurun_out/bench_refstyle.json'));print('refstyle', d['value']/1e6, d['e2e']['value']/1e6, d['checks'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --units 4000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --fastq-units 0 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fp_chain -s 1 -c 1 -o gpurun_out/prof_chain \
    python bench.py --units 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --fastq-units 0 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -30
