#!/bin/bash
# text-path check: FASTQ codec + host shim parity tests, then the bench line with its fastq_path object
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 -k "fastq or shim" 2>&1 | tail -4
python bench.py --units 8000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/b.err
python -c "
import json;d=json.load(open('gpurun_out/b.json'));fq=d['fastq_path']
print('VALUE %.1f M/s' % (d['value']/1e6)); print('FASTQ', {k:(round(v/1e6,2) if k=='value' else v) for k,v in fq.items() if k in ('value','two_workers','decode','encode','error')})"
tail -2 gpurun_out/b.err
