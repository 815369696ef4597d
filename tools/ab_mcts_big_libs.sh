#!/bin/bash
# A/B of library variants in the config-5 regime (8192 games per GPU, the 28-per-SM one-warp kernel): opening plies + complete shard
for v in "$@"; do
  f=blokus-engine_b200/lib/libblokus_b200$v.so
  [ "$v" == "default" ] && f=blokus-engine_b200/lib/libblokus_b200.so
  echo "== variant '$v'"
  BK_LIB=$f BK_BIG=1 python tools/probe_mcts.py
  BK_LIB=$f WORLD_SIZE=8 RANK=0 python tools/bench_config5.py | cut -c100-330
done
