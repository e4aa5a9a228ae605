#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "stw" 2>&1 | tail -3
tools/_exp.sh
