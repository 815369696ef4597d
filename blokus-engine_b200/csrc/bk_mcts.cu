// bk_mcts.cu — MCTS self-play clients (placeholder entry points; implementation in progress).
#include "bk_host.h"

struct bk_selfplay { int n; };

extern "C" {
#define BK_NOT_YET(name) return bk_fail(BK_ERR_STATE, name ": not implemented yet")
int bk_selfplay_create(int, int, const bk_config*, uint32_t, uint32_t, bk_selfplay**) { BK_NOT_YET("bk_selfplay_create"); }
void bk_selfplay_destroy(bk_selfplay*) {}
int bk_selfplay_run_stub(bk_selfplay*, int) { BK_NOT_YET("bk_selfplay_run_stub"); }
int bk_selfplay_begin_ply(bk_selfplay*) { BK_NOT_YET("bk_selfplay_begin_ply"); }
int bk_selfplay_leaf_planes(bk_selfplay*, float*, int32_t*) { BK_NOT_YET("bk_selfplay_leaf_planes"); }
int bk_selfplay_expand_backup(bk_selfplay*, const float*, const float*, int32_t*) { BK_NOT_YET("bk_selfplay_expand_backup"); }
int bk_selfplay_end_ply(bk_selfplay*) { BK_NOT_YET("bk_selfplay_end_ply"); }
int bk_selfplay_live_games(bk_selfplay*, int32_t*) { BK_NOT_YET("bk_selfplay_live_games"); }
bk_env* bk_selfplay_env(bk_selfplay*) { return nullptr; }
int bk_selfplay_results(bk_selfplay*, int32_t*, int32_t*, int32_t, int16_t*, uint32_t*) { BK_NOT_YET("bk_selfplay_results"); }
int bk_selfplay_last_root(bk_selfplay*, int32_t*, int16_t*, uint32_t*, float*, float*) { BK_NOT_YET("bk_selfplay_last_root"); }
int bk_selfplay_counters(bk_selfplay*, uint64_t*) { BK_NOT_YET("bk_selfplay_counters"); }
int bk_selfplay_last_kernel_ms(bk_selfplay*, float*) { BK_NOT_YET("bk_selfplay_last_kernel_ms"); }
}
