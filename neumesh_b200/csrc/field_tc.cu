// tcgen05 MLP engine (placeholder until the tensor-core kernels land in this file).
#include "field.cuh"

namespace nmb {
int pack_mlp_tc(const nmb_field_desc*, const FieldLayout&, nmb_field*, cudaStream_t) { return 0; }
int launch_geo_tc(const nmb_field*, const FieldIn&, int64_t, float*, float*, cudaStream_t) {
  set_error("tcgen05 MLP engine not built");
  return 4;
}
int launch_color_tc(const nmb_field*, const FieldIn&, int64_t, float*, cudaStream_t) {
  set_error("tcgen05 MLP engine not built");
  return 4;
}
}  // namespace nmb
