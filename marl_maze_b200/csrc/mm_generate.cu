// mm_generate.cu -- K1: batched maze generator (placeholder until the carve kernel lands in this file).
#include "mm_env.cuh"
namespace mm {
cudaError_t launch_generate(const mm_state*, int, int, int, int, int, int, uint64_t, uint32_t, void*, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace mm
