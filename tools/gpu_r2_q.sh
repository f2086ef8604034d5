#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_autograd.py -m gpu -x -q 2>&1 | tail -15
