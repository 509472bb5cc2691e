#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "big" 2>&1 | tail -30
