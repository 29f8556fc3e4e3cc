// mm_update.cuh -- argument block of the heads + PPO-loss kernel (mm_update.cu) and the launchers of the K5 update kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mm {

struct PpoLossArgs {
    const float *h2, *head_w, *head_b;
    const uint8_t *masks, *actions;
    const float *old_logp, *adv;
    float clip, scale;
    float *dz2, *logp, *part;
    int E;
};

int ppo_loss_blocks();
int ppo_loss_part_ld();
cudaError_t launch_ppo_heads_loss(const PpoLossArgs& a, cudaStream_t stream);
cudaError_t launch_wgrad_tc(const float* dz, const float* h, int R, int n_out, int k_in, float* part, cudaStream_t stream);
void wgrad_geometry(int R, int n_out, int k_in, int* slabs, int* ld, int* kb_per, int* out_rows, int* transposed);
int segment_sum_blocks(int rows);
cudaError_t launch_segment_sum(const float* x, const long long* seg, int rows, int cols, int n_seg, float* part, cudaStream_t stream);
cudaError_t launch_gather_rows(const float* src, const long long* seg, int rows, int cols, int n_src, float* out, cudaStream_t stream);
struct HeadArgs;
cudaError_t launch_linear_tc_ex(const float* x, const float* w_hi, const float* w_lo, int n_rows_w, const float* bias, float* y, int ldy, int M, int K, int mode,
                                const uint32_t* gate, uint32_t* gate_out, const float* head_w, const float* head_b, const HeadArgs* heads, cudaStream_t stream);

}  // namespace mm
