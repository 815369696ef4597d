#!/bin/bash
# 28 vs 32 resident games per SM with in-place candidate tables (profiles/r02_ab_mcts_stub_global_cands2.log); the 32-per-SM
# instantiation (64 registers, 56 B of spills, 3 % slower) was removed afterwards, so BK_STUB_MIN_BLOCKS=32 now runs the 28 build.
for mb in 28 32; do
  echo "== minb=$mb  (8192 games x 12 opening plies)"
  BK_STUB_MIN_BLOCKS=$mb BK_BIG=1 python tools/probe_mcts.py
  echo "== minb=$mb config-5 shard"
  BK_STUB_MIN_BLOCKS=$mb WORLD_SIZE=8 RANK=0 python tools/bench_config5.py
done
echo "== pipe (global cands) 1024 complete games"
BK_FULLGAME=1 python tools/probe_mcts.py
