"""fastp_b200 -- B200-native per-read FASTQ preprocessing hot path (drop-in for fastp's worker body).

The product is fastp_b200/libfastp_b200.so (CUDA sm_100a kernels + C-ABI, include/fastp_b200.h);
this package only holds the ctypes mirror used by tests and bench.py.
"""
from . import capi  # noqa: F401
