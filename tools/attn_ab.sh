#!/bin/bash
# A/B of the self-attention kernels on one box, back to back (each configuration is its own process: the switches are read once).
out=${1:-gpurun_out/attn_ab.log}
: > "$out"
for shape in "16 4096 8 40" "8 9216 8 40" "16 1024 8 80" "16 256 8 160" "8 257 16 64"; do
  for cfg in "PBE_ATTN_KERNEL=2" "PBE_ATTN_POLY=0" "PBE_ATTN_POLY=2" "PBE_ATTN_POLY=4" "PBE_ATTN_POLY=6" "PBE_ATTN_POLY=8" "PBE_ATTN_POLY=4 PBE_ATTN_RERUN=0" "PBE_ATTN_POLY=4 PBE_ATTN_TWO_PASS=1"; do
    case "$shape" in *" 80"|*" 160") case "$cfg" in "PBE_ATTN_POLY=0"|"PBE_ATTN_POLY=4 PBE_ATTN_RERUN=0"|"PBE_ATTN_POLY=4 PBE_ATTN_TWO_PASS=1") ;; *) continue;; esac;; esac
    env $cfg timeout 120 python tools/attn_probe.py $shape 2>&1 | tail -1 >> "$out"
  done
done
cat "$out"
