#!/bin/bash
# tools/ncu_export.sh gpurun_out/NAME   (on the GPU box, after `ncu -o gpurun_out/NAME ...`):
# export the raw page and the per-line source page as CSV and drop the 25 MB report, so that several captures fit into
# the 64 MiB that travel back.  Read them here with tools/ncu_figures.py NAME_raw.csv / tools/ncu_lines.py NAME_src.csv.gz
for base in "$@"; do
  rep=$base.ncu-rep
  [ -f "$rep" ] || { echo "no $rep"; continue; }
  ncu -i "$rep" --page raw --csv > "${base}_raw.csv" 2>/dev/null
  ncu -i "$rep" --page source --print-source cuda,sass --csv 2>/dev/null | gzip > "${base}_src.csv.gz"
  rm -f "$rep"
done
