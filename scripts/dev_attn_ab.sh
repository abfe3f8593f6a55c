#!/bin/bash
# interleaved A/B of the 64x64-level reference attention variants (separate processes; read the min per variant)
for r in 1 2 3; do
  echo "round $r"
  AP_ATTENTION_V5=1 timeout 200 python scripts/dev_attn_time.py 2>&1 | tail -1 | sed 's/^/denom      /'
  AP_ATTENTION_NO_DENOM=1 timeout 200 python scripts/dev_attn_time.py 2>&1 | tail -1 | sed 's/^/no_denom   /'
  AP_ATTENTION_V5=0 timeout 200 python scripts/dev_attn_time.py 2>&1 | tail -1 | sed 's/^/v3         /'
done
