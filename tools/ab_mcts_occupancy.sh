#!/bin/bash
# stub kernel occupancy variants chosen through BK_STUB_MIN_BLOCKS (0 = the library's own choice)
for mb in 0 1 12 16 20; do   # only these builds exist; 0 = the library picks by batch size
  echo "== BK_STUB_MIN_BLOCKS=$mb"
  BK_STUB_MIN_BLOCKS=$mb BK_FULLGAME=1 python tools/probe_mcts.py
  BK_STUB_MIN_BLOCKS=$mb BK_BIG=1 python tools/probe_mcts.py
done
