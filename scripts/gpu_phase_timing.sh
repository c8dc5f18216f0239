#!/bin/bash
# phase timing on the GPU box: swaps in the -DFP_PHASE_TIMING build (scripts/timing/libfastp_b200.so), prints what the 16 warps of CTA 0 spent per phase
mkdir -p gpurun_out
[ -f scripts/timing/libfastp_b200.so ] || { echo "build it first: nvcc ... -DFP_PHASE_TIMING ... -o scripts/timing/libfastp_b200.so (same line as __graft_entry__.build)"; exit 1; }
cp fastp_b200/libfastp_b200.so /tmp/lib_orig.so; cp scripts/timing/libfastp_b200.so fastp_b200/libfastp_b200.so
for wl in ${WLS:-pe150_overlap_correction}; do
  python bench.py --workload $wl --units ${U:-8000000} --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --fastq-units 0 --no-workloads > gpurun_out/timing_$wl.log 2>&1
  grep PHASE gpurun_out/timing_$wl.log | tail -16
done
cp /tmp/lib_orig.so fastp_b200/libfastp_b200.so
