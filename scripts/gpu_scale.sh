#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -s 2>&1 | tail -15
