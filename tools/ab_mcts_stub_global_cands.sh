#!/bin/bash
# Record of the A/B that made in-place candidate tables the default (profiles/r02_ab_mcts_stub_global_cands.log):
# '' = the library with the 5 KB shared-memory copy per game, '_sgc' = the same sources built with the (since removed)
# -DBK_STUB_GLOBAL_CANDS switch, i.e. today's default.  To repeat it, build the old side from commit 14cbde9.
for v in "" _sgc; do
 f=blokus-engine_b200/lib/libblokus_b200$v.so
 for mb in 20 28; do
  echo "== lib '$v' minb=$mb  (8192 games x 12 opening plies)"
  BK_LIB=$f BK_STUB_MIN_BLOCKS=$mb BK_BIG=1 python tools/probe_mcts.py
 done
 echo "== lib '$v' config-5 shard (library's own choice)"
 BK_LIB=$f WORLD_SIZE=8 RANK=0 python tools/bench_config5.py
done
echo "== pipe default (global cands now) 1024 complete games"
BK_FULLGAME=1 python tools/probe_mcts.py
