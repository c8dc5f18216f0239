#!/bin/bash
# round-end check on the GPU box: the chain-kernel parity files, then the default bench line
mkdir -p gpurun_out
timeout ${TEST_TIMEOUT:-800} python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_host_shim.py tests/test_gpu_reference_binding.py -m gpu -q -x --timeout 600 2>&1 | tail -6 > gpurun_out/final_pytest.log; cat gpurun_out/final_pytest.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 2500 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
