#!/bin/bash
# A/B of library variants on the config-3 shape (complete games, 1024 games, 800 sims/move)
mkdir -p gpurun_out
L=blokus-engine_b200/lib
for v in "" $@; do
  f=$L/libblokus_b200$v.so
  echo "== variant '$v'" 
  BK_LIB=$f BK_FULLGAME=1 python tools/probe_mcts.py
done
