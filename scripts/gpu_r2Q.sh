#!/bin/bash
# round 2: cdf_kernel with fewer CTAs in flight (live workspaces vs the 126 MB L2)
for g in 296 222 148 111 74 48; do
  echo "ctas=$g $(PWMSCAN_CDF_CTAS=$g timeout 300 python scripts/probe_cdf.py 800 2>&1 | tail -1)"
done
