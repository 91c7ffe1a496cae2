#!/bin/bash
# Turns gpurun_out/<name>.ncu-rep into <name>.raw.csv (+ <name>.source.csv) and removes the report: gpurun only
# brings back 64 MiB, and a --set full report with sources is 20-35 MB.
for rep in "$@"; do
  base="${rep%.ncu-rep}"
  ncu -i "$rep" --page raw --csv > "$base.raw.csv" 2>/dev/null
  ncu -i "$rep" --page source --csv > "$base.source.csv" 2>/dev/null
  rm -f "$rep"
  gzip -f "$base.source.csv"
done
