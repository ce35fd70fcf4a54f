#!/bin/bash
# Development aid: sweep the tile configurations of the int8 class-sum kernel on the C2 shape.
for v in 0 1 2; do
  echo "variant $v: $(GWSEM_IMMA_VARIANT=$v timeout 200 python scripts/quick_time.py 100000 65536 2>&1 | grep per-kernel)"
done
