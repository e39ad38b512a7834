#!/bin/bash
# Attention ablation sweep (micro-benchmark only; B200PF_ATTN_DBG bits: 1 no exponentials, 4 S with one K step, 8 no P V,
# 16 softmax warps idle, 32 odd query tiles skip their K/V loads; B200PF_ATTN_CTAS = resident CTAs per SM).
#   bash tools/ablate_attn.sh > gpurun_out/attn_ablation.txt
for C in 2 1; do
  for D in 0 1 4 8 16 20 24 28 32 48; do
    echo "== ctas_per_sm=$C dbg=$D"
    B200PF_ATTN_CTAS=$C B200PF_ATTN_DBG=$D python tools/bench_attn.py 1024 10 2>&1 | tail -4
  done
done
