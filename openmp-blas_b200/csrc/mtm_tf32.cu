// Placeholder until the tcgen05 3xTF32 path lands: reports "no configs" so AUTO stays on SIMT.
#include "mtm_kernels.h"

namespace b200 {
namespace {
const TileConfig kNone = {"", 0, 0, 0, 0, 0};
}
int tf32_num_configs() { return 0; }
const TileConfig& tf32_config(int) { return kNone; }
size_t tf32_workspace_bytes(const MtmShape&) { return 0; }
cudaError_t launch_3xtf32_f32(float*, const float*, const float*, const MtmShape&, void*, size_t, int,
                              cudaStream_t, int*) {
    return cudaErrorNotSupported;
}
}  // namespace b200
