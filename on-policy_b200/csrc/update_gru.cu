// update_gru.cu -- training step for recurrent (GRU) nets.  (filled in below the MLP path)
#include "net_tiles.cuh"

namespace mappo {

int update_gru_slots(const NetDev&, int, int, int) { return 1; }
int64_t update_gru_workspace_floats(const NetDev&, int) { return 0; }
int update_gru_launch(const NetDev&, const float*, const BatchDev&, const LossDev&, const double*, const double*,
                      const float*, float*, int, double*, float*, cudaStream_t) {
  set_error("update_fwd_bwd: recurrent nets are not built yet in this library version");
  return MAPPO_ERR_UNSUPPORTED;
}

}  // namespace mappo
