// placeholder until the tcgen05 kernel lands
#include "match_tc.cuh"
namespace dsir {
bool match_tc_supported(const dsir_feat &, const dsir_feat &, int, int, int, int) { return false; }
bool match_tc_profitable(int, int, int, int) { return false; }
size_t match_tc_workspace_bytes(int, int, int, int) { return 0; }
int launch_match_tc(const MatchParams &, void *, size_t, cudaStream_t) { return DSIR_ERR_UNSUPPORTED; }
}  // namespace dsir
