#!/bin/bash
# Development aid: counts of the SASS mnemonics that matter (tensor-core, TMA, FP64, staging) per object file.
cd "$(dirname "$0")/../gwsem_b200/csrc" || exit 1
for f in stats_tc5 marg_series marg_fit stats_imma stats marg_stats_j3 marg_exact fit ml_fit records decode; do
  echo "== $f.o"
  cuobjdump -sass $f.o 2>/dev/null | grep -oE '^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+' | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn \
    | awk '$2 ~ /^(UTCIMMA|STTM|LDTM|UTMALDG|UBLKCP|IMMA|DFMA|DMUL|DADD|LDGSTS|MUFU|BAR|LDS|STS|LDG|STG|SYNCS|UTCBAR|UTCCP|PRMT|POPC)$/ {printf "%s:%s ", $2, $1} END {print ""}'
done
