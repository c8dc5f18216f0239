export TESTS='tests/test_gpu_parity.py tests/test_gpu_golden.py'
export TESTK="${TESTK:-enriched_device_mode or random_option or patch or ref_style or golden or edge or non_acgtn or len250}"
VARIANTS="${VARIANTS:-FP_XFLAGS=2}" WLS="${WLS:-pe150_overlap_correction pe150_full se150_cut_right_polyg pe250_overrep}" bash scripts/gpu_ab.sh
