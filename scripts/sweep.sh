#!/bin/bash
# development aid: sweep class_sums tuning variants on the C2 shape
for v in 0 1 2 3 4 5 6 7; do
  echo "variant $v"; GWSEM_CS_VARIANT=$v python scripts/quick_time.py 100000 28416 item 2>&1 | sed -n 3p
done
