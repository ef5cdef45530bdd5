// Host-mode builds of the P2P engine (sanitizer runs without nvcc): the copy kernel is never launched when the
// endpoint has no GPU, only its launcher symbols are needed.
#include <cuda_runtime.h>

#include "p2p_types.h"

namespace ub {
cudaError_t launch_p2p_copy(const P2PCopyBatch&, int, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t preload_p2p_kernels() { return cudaSuccess; }
}  // namespace ub
