#!/bin/bash
# A/B of library variants (blokus-engine_b200/lib/libblokus_b200<suffix>.so) on the config-3 shape: complete games, 1024 games
for v in "$@"; do
  f=blokus-engine_b200/lib/libblokus_b200$v.so
  [ "$v" == "default" ] && f=blokus-engine_b200/lib/libblokus_b200.so
  echo "== variant '$v'"
  BK_LIB=$f BK_FULLGAME=1 python tools/probe_mcts.py
done
