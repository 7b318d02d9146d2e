#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -30 > gpurun_out/r11_pytest.log; tail -25 gpurun_out/r11_pytest.log
