export TESTS='tests/test_gpu_parity.py tests/test_gpu_golden.py'
export TESTK='enriched_device_mode or random_option or patch or ref_style or golden or edge or non_acgtn or len250 or merge_mode'
VARIANTS='FP_XFLAGS=2' WLS='pe150_overlap_correction' bash scripts/gpu_ab.sh
FP_XFLAGS=3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "random_option or cfg3 or patch" 2>&1 | tail -3
