for mb in 28 32; do
  echo "== minb=$mb  (8192 games x 12 opening plies)"
  BK_STUB_MIN_BLOCKS=$mb BK_BIG=1 python tools/probe_mcts.py
  echo "== minb=$mb config-5 shard"
  BK_STUB_MIN_BLOCKS=$mb WORLD_SIZE=8 RANK=0 python tools/bench_config5.py
done
echo "== pipe (global cands) 1024 complete games"
BK_FULLGAME=1 python tools/probe_mcts.py
