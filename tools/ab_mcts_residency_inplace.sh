#!/bin/bash
# which instantiation of the one-warp kernel for mid-size batches, now that shared memory no longer caps residency
for n in 1536 2048 3072 4096; do
 for mb in 0 12 16 20 28; do
  echo "== n=$n BK_STUB_MIN_BLOCKS=$mb"
  BK_N=$n BK_PLIES=12 BK_STUB_MIN_BLOCKS=$mb python tools/probe_mcts.py
 done
done
echo "== row f3 modes at 8192 games (28-per-SM instantiation with the opt-in modes)"
python tools/probe_modes.py 8192
