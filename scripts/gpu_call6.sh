#!/bin/bash
mkdir -p gpurun_out
EVERY=4 timeout 300 python scripts/diag_transient.py 100 "screen_rounds_busy=-1" > gpurun_out/r06_trans_off.txt 2>&1
EVERY=4 timeout 300 python scripts/diag_transient.py 100 "" > gpurun_out/r06_trans_adaptive.txt 2>&1
paste -d'\n' gpurun_out/r06_trans_off.txt gpurun_out/r06_trans_adaptive.txt | tail -56
