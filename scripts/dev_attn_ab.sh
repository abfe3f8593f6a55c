#!/bin/bash
# interleaved A/B of the 64x64-level reference attention (attention5_kernel) start-up stagger of the SM's second CTA
# (separate processes: the library reads AP_ATTN_STAGGER once; read the min per variant)
for r in 1 2; do
  echo "round $r"
  for st in 0 400 700 1000 1400; do
    AP_ATTN_STAGGER=$st timeout 200 python scripts/dev_attn_time.py 2>&1 | tail -1 | sed "s/^/stagger=$st  /"
  done
done
