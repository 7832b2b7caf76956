// Library-level entry points: version, last CUDA error text.
#include "common.cuh"

namespace mma {
static thread_local cudaError_t g_last = cudaSuccess;
void set_last_error(cudaError_t e) { g_last = e; }
}  // namespace mma

extern "C" int mma_b200_version(void) { return 100; }
extern "C" const char *mma_last_cuda_error(void) { return cudaGetErrorString(mma::g_last); }
