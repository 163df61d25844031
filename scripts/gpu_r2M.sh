#!/bin/bash
# round 2: which entry point makes the process abort at exit under MALLOC_PERTURB_ / MALLOC_CHECK_?
for s in nothing scan scan_keep iupac scores max cdfs affinity many; do
  MALLOC_CHECK_=3 MALLOC_PERTURB_=165 timeout 120 python scripts/probe_exit.py $s > /tmp/pe_$s.log 2>&1; echo "$s rc=$? $(tail -1 /tmp/pe_$s.log | cut -c1-100)"
done
