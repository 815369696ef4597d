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
