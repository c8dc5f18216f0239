#!/bin/bash
# quick look at every bench workload (8 M units each) + the text path with two workers
mkdir -p gpurun_out
for W in pe150_overlap_correction pe150_full se150_cut_right_polyg; do
  python bench.py --workload $W --units 8000000 --steps 3 --warmup 3 --no-cpu-baseline --fastq-units ${FQ:-0} > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_$W.json'))
print('$W VALUE %.1f M/s kernel_ms %.2f e2e %.1f M/s' % (d['value']/1e6, d['roofline']['kernel_ms'], d.get('e2e',{}).get('value',0)/1e6), d['checks'])
fq=d.get('fastq_path')
if fq: print('  FASTQ', {k:(round(v/1e6,2) if k=='value' else v) for k,v in fq.items() if k in ('value','two_workers','decode','encode','cpu_cli','error')})"
  tail -2 gpurun_out/bench_$W.err
done
python bench.py --profile 0 --units 8000000 --steps 3 --warmup 3 --no-cpu-baseline --fastq-units 0 > gpurun_out/bench_refstyle.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_refstyle.json'));print('refstyle VALUE %.1f M/s' % (d['value']/1e6))"
