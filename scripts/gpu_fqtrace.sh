#!/bin/bash
mkdir -p gpurun_out
FP_FQ_TRACE=1 python bench.py --units 2000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/t.json 2> gpurun_out/t.err
grep "\[fq\]" gpurun_out/t.err | tail -24
python -c "
import json;d=json.load(open('gpurun_out/t.json'));fq=d['fastq_path'];print(fq.get('value'), fq.get('two_workers'), fq.get('error'))"
