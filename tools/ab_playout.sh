#!/bin/bash
# A/B of library variants (blokus-engine_b200/lib/libblokus_b200<suffix>.so) on config 2's kernel: k_playout, 4096 / 16384 / 65536 games
for v in "$@"; do
  f=blokus-engine_b200/lib/libblokus_b200$v.so
  [ "$v" == "default" ] && f=blokus-engine_b200/lib/libblokus_b200.so
  echo "== variant '$v'"
  BK_LIB=$f python tools/probe_playout.py
done
