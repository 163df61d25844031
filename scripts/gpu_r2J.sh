#!/bin/bash
# round 2: the packed store on disk — scanGenome(packed=) and the app tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_apps.py -m gpu -x -q 2>&1 | tail -3
