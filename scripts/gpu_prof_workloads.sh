#!/bin/bash
# ncu --set full capture of the chain kernel for each BASELINE workload (second launch = a timed-like step), reports into gpurun_out/
mkdir -p gpurun_out
for wl in ${WLS:-pe150_full se150_cut_right_polyg pe250_overrep}; do
  ncu --set full --clock-control none --import-source on -k regex:fp_chain -s 1 -c 1 -o gpurun_out/prof_$wl -f \
    python bench.py --workload $wl --units 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --fastq-units 0 --no-workloads > gpurun_out/ncu_$wl.log 2>&1
  ls -la gpurun_out/prof_$wl.ncu-rep
done
