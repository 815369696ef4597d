#!/bin/bash
# (1) two-warp pipeline vs one-warp kernel just above the pipeline's residency (8 CTAs per SM = 1184 games on 148 SMs)
# (2) which one-warp instantiation for small batches when the pipeline does not apply (opt-in modes): all registers (<1>) or <12>
for n in 296 592 1024 1184 1300; do
  for v in "BK_STUB_MIN_BLOCKS=1" "BK_STUB_MIN_BLOCKS=12" "BK_STUB_PIPE=1" "BK_X=0"; do
    echo "== n=$n $v"
    env $v BK_N=$n BK_PLIES=12 python tools/probe_mcts.py | head -1
  done
done
