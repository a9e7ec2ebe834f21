// placeholder until the bitmap k-clique kernel lands (next commit)
#include "gm_internal.cuh"
namespace gm {
int prepare_kclique_bitmap(gm_graph *g) { return ensure_coo(g, 0); }
int run_kclique_bitmap(gm_graph *, int, int *, bool *handled) { *handled = false; return GM_OK; }
}
