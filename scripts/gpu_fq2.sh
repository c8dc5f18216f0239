#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 -k "fastq or shim" 2>&1 | tail -4
FP_FQ_TRACE=1 python bench.py --units 2000000 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/t.json 2> gpurun_out/t.err
grep "\[fq\]" gpurun_out/t.err | sed -n 1,9p
python -c "
import json;d=json.load(open('gpurun_out/t.json'));fq=d['fastq_path'];print(fq.get('value'), fq.get('two_workers'), fq.get('decode'), fq.get('error'))"
