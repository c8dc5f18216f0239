#!/usr/bin/env python
"""md5 of the UNMODIFIED reference CLI's output FASTQ files (oracle/_ref/fastp_ref -w 1) for the text-in / text-out
cases of tests/test_gpu_fastq.py -> tests/golden/fastq_cli_digests.json (for boxes without the reference binary)."""
import hashlib
import json
import os
import sys
import tempfile
from pathlib import Path

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_fastq as G  # noqa: E402

out = {}
for case in G.CLI_CASES:
    for paired in (1, 0):
        flags, p, t1, t2 = G.cli_inputs(case, paired)
        with tempfile.TemporaryDirectory() as d:
            outs = G.run_cli(Path(d), flags, t1, t2)
        out[f"{case}/{'pe' if paired else 'se'}"] = [hashlib.md5(x).hexdigest() for x in outs]
json.dump(out, open(os.path.join(HERE, "fastq_cli_digests.json"), "w"), indent=1)
print(out)
