#!/bin/bash
# One GPU round trip while tuning the narrow beam kernel: fast parity subset, then per-phase cycles and
# the batch sweep.   tools/perf_check.sh <tag>   -> gpurun_out/<tag>.txt
tag=${1:-perf}
mkdir -p gpurun_out
{
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast" 2>&1 | tail -3
  python tools/phase_cycles.py 148 gauss
  python tools/phase_cycles.py 148 peaky
  python tools/batch_sweep.py gauss 148 256 592 1024 8192
  python tools/batch_sweep.py peaky 256 1024
} > gpurun_out/$tag.txt 2>&1
grep -v "^$" gpurun_out/$tag.txt | tail -14
