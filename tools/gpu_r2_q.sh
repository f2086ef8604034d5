#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_autograd.py -m gpu -x -q -k "zero_extent or default_optimizer or autograd or gradients" 2>&1 | tail -8
