#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --tb=short -k "fp32" 2>&1 | grep -v "^  \|^$" | tail -20
timeout 1500 python scripts/sweep.py r02 > gpurun_out/r16_sweep.log 2>&1; tail -5 gpurun_out/r16_sweep.log; cp profiles/r02_sweep.* gpurun_out/ 2>/dev/null; cat profiles/r02_sweep.md 2>/dev/null | head -40
