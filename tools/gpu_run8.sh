#!/bin/bash
set -x
mkdir -p gpurun_out
for l in 557fd44 4331b6a 1332d6c head; do timeout 120 python tools/bisect_stem.py tools/_bisect/lib_$l.so; done > gpurun_out/run8_bisect.log 2>&1
ODTK_STEM_ROWS=0 timeout 120 python tools/bisect_stem.py tools/_bisect/lib_head.so >> gpurun_out/run8_bisect.log 2>&1
ODTK_STEM_RAW=0 timeout 120 python tools/bisect_stem.py tools/_bisect/lib_head.so >> gpurun_out/run8_bisect.log 2>&1
cat gpurun_out/run8_bisect.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/run8_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/run8_pytest.log
grep -E "passed|failed|^FAILED|rc=" gpurun_out/run8_pytest.log | head -30
