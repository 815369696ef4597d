#!/bin/bash
for mb in 20 24 32; do
  echo "== BK_STUB_MIN_BLOCKS=$mb"
  BK_STUB_MIN_BLOCKS=$mb BK_BIG=1 python tools/probe_mcts.py
done
