#!/bin/bash
# A/B of library variants: config-3 shape complete games at 1024 games, and 12 plies at 8192 games
L=blokus-engine_b200/lib
for v in "" $@; do
  f=$L/libblokus_b200$v.so
  echo "== variant '$v'"
  BK_LIB=$f BK_FULLGAME=1 python tools/probe_mcts.py
  BK_LIB=$f BK_BIG=1 python tools/probe_mcts.py
done
