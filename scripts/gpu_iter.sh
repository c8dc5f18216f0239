#!/bin/bash
# iteration loop on the GPU box: parity tests, then a short bench line (and optionally an ncu capture)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
grep -q "passed" gpurun_out/pytest_gpu.log && ! grep -q "failed" gpurun_out/pytest_gpu.log || exit 1
U=${U:-8000000}
python bench.py --units $U --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python -c "
import json;d=json.load(open('gpurun_out/bench_iter.json'))
print('VALUE %.1f M/s  kernel_ms %.2f  frac %.4f  e2e %.1f M/s' % (d['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('e2e',{}).get('value',0)/1e6), d['checks'], d['clocks']); print('FASTQ', json.dumps(d.get('fastq_path')))"
tail -3 gpurun_out/bench_iter.err
if [ -n "$NCU" ]; then
  ncu --set full --clock-control none --import-source on -k regex:fp_chain -s 1 -c 1 -o gpurun_out/prof_iter \
    python bench.py --units 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e ${BENCH_ARGS} > gpurun_out/ncu_iter.log 2>&1
  ls -la gpurun_out/prof_iter.ncu-rep
fi
