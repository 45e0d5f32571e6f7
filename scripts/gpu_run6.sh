#!/bin/bash
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "infonce" 2>&1 | tail -15
timeout 600 python scripts/bench_infonce.py 2>&1 | tail -8
