#!/bin/bash
# interleaved A/B of the 64x64-level reference attention generations (separate processes; read the min per variant)
for r in 1 2 3; do
  echo "round $r"
  timeout 200 python scripts/dev_attn_time.py 2>&1 | tail -1 | sed 's/^/v5  /'
  AP_ATTENTION_V5=0 timeout 200 python scripts/dev_attn_time.py 2>&1 | tail -1 | sed 's/^/v3  /'
done
