#!/bin/bash
MALLOC_CHECK_=3 MALLOC_PERTURB_=165 PYTHONFAULTHANDLER=1 timeout 300 python scripts/probe_aux.py > /tmp/pa.log 2>&1; echo "rc=$?"; tail -4 /tmp/pa.log | cut -c1-160
