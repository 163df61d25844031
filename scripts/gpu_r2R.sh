#!/bin/bash
# round 2: maxMatchSc with look-ahead pruning behind a first stretch
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q -k "dense or max_match" 2>&1 | tail -3
timeout 300 python scripts/probe_aux.py 2>&1 | grep maxMatchSc | cut -c1-200
