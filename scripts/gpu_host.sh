#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_gpu.py -m gpu -q 2>&1 | tail -40
