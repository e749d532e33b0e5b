// Internal interface of the tcgen05 weight-gradient kernel (wgrad_tc.cu), used by eda_wgrad (grad_ops.cu).
#pragma once
#include <cuda_runtime.h>
#include "../../include/eda_b200.h"

namespace eda {
// Shapes / alignments the tensor-core kernel takes (everything else stays on the warp-level kernel).
bool wgrad_tc_eligible(const eda_wgrad_problem *probs, int nprobs, int N, int K);
// Launches it; kWgradTcDeclined = nothing launched (tensor maps could not be encoded), else an EDA_* code.
constexpr int kWgradTcDeclined = -1000;
int wgrad_tc_timestamps(long long *host_out, int n);  // development aid, see wgrad_tc.cu
int wgrad_tc_launch(const eda_wgrad_problem *probs, int nprobs, int N, int K, cudaStream_t stream);
}  // namespace eda
