// placeholder until the P2P engine lands
#include "../kernels/launch.h"
namespace ub {
cudaError_t preload_p2p_kernels() { return cudaSuccess; }
}  // namespace ub
