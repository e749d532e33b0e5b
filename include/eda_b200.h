/*
 * eda_b200 — C ABI of the B200-native (sm_100a) hot path of yanmin-wu/EDA.
 *
 * This header is the drop-in boundary.  Every entry point takes plain device pointers,
 * sizes and a CUDA stream handle (cudaStream_t passed as void*; NULL = legacy default
 * stream) and returns 0 on success or a negative EDA_ERR_* code.  Nothing here exits the
 * process (the reference's CUDA_CHECK_ERRORS() calls exit(-1),
 * pointnet2/_ext_src/include/cuda_utils.h:35-44), nothing allocates: the CALLER owns every
 * buffer including scratch, and every call is asynchronous with respect to the host and
 * ordered on `stream` — the same stream contract as the reference ops, which launch on
 * at::cuda::getCurrentCUDAStream() of the current device.
 *
 * All tensors are dense row-major ("contiguous"); float = IEEE fp32, index = int32,
 * exactly as the reference checks (pointnet2/_ext_src/include/utils.h:10-30).
 *
 * Each function cites the reference interface it replaces (path relative to the
 * reference repository root).  The reference-side binding a maintainer adds is shown in
 * INTEGRATION.md.
 */
#ifndef EDA_B200_H_
#define EDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define EDA_API __attribute__((visibility("default")))
#else
#define EDA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define EDA_OK 0
#define EDA_ERR_INVALID_ARGUMENT (-1) /* null pointer, negative size, unsupported shape */
#define EDA_ERR_CUDA_LAUNCH (-2)      /* cudaGetLastError() != cudaSuccess after launch */
#define EDA_ERR_NO_DEVICE (-3)        /* no sm_100 device / driver */
#define EDA_ERR_UNSUPPORTED (-4)      /* configuration outside the compiled template set */

/* Library/ABI version (major*10000 + minor*100 + patch). */
EDA_API int eda_version(void);
/* Static human-readable string for an EDA_ERR_* code (never NULL). */
EDA_API const char *eda_error_string(int code);
/* Text of the last CUDA error seen by this thread inside the library ("" if none). */
EDA_API const char *eda_last_cuda_error(void);

/* ---------------------------------------------------------------------------------------
 * Furthest point sampling.
 * Replaces: at::Tensor furthest_point_sampling(at::Tensor points, const int nsamples)
 *           pointnet2/_ext_src/src/sampling.cpp:70-91 (kernel sampling_gpu.cu:74-178).
 * xyz (B,N,3) f32 -> idxs (B,m) i32, bit-exact with the reference incl. its tie-break.
 * `scratch`: device buffer of eda_fps_scratch_bytes(B,N,m) bytes (may be NULL when that
 * is 0).  The reference allocates its (B,N) `tmp` per call (sampling.cpp:78-80); here the
 * running minima live in registers and scratch is only needed by the large-N fallback.
 */
EDA_API size_t eda_fps_scratch_bytes(int B, int N, int m);
EDA_API int eda_furthest_point_sampling(const float *xyz, int B, int N, int m, void *scratch, int *idxs,
                                void *stream);

/* Ball query.
 * Replaces: at::Tensor ball_query(at::Tensor new_xyz, at::Tensor xyz, const float radius,
 *           const int nsample)  pointnet2/_ext_src/src/ball_query.cpp:13-37
 *           (kernel ball_query_gpu.cu:14-49).
 * new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample) i32; every slot is written (zeros for
 * an empty ball), so idx need not be pre-zeroed.  Bit-exact. */
EDA_API int eda_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M, float radius,
                   int nsample, int *idx, void *stream);

/* Grouping.  Replaces group_points / group_points_grad,
 * pointnet2/_ext_src/src/group_points.cpp:17-65 (kernels group_points_gpu.cu:13-80).
 * points (B,C,N), idx (B,M,S) -> out (B,C,M,S).
 * grad: grad_out (B,C,M,S) -> grad_points (B,C,N); the callee zero-fills grad_points. */
EDA_API int eda_group_points(const float *points, const int *idx, int B, int C, int N, int M, int S,
                     float *out, void *stream);
EDA_API int eda_group_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int M, int S,
                          float *grad_points, void *stream);

/* Gather.  Replaces gather_points / gather_points_grad,
 * pointnet2/_ext_src/src/sampling.cpp:20-69 (kernels sampling_gpu.cu:13-62).
 * points (B,C,N), idx (B,M) -> out (B,C,M);  grad_out (B,C,M) -> grad_points (B,C,N) (zero-filled here). */
EDA_API int eda_gather_points(const float *points, const int *idx, int B, int C, int N, int M, float *out,
                      void *stream);
EDA_API int eda_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int M,
                           float *grad_points, void *stream);

/* 3-NN + interpolation.  Replaces three_nn / three_interpolate / three_interpolate_grad,
 * pointnet2/_ext_src/src/interpolate.cpp:19-104 (kernels interpolate_gpu.cu:14-159).
 * unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3) f32 (SQUARED distances, as the reference
 * op returns; the sqrt is applied by the Python wrapper, pointnet2_utils.py:142), idx (B,n,3).
 * points (B,C,m), idx (B,n,3), weight (B,n,3) -> out (B,C,n).
 * grad_out (B,C,n) -> grad_points (B,C,m) (zero-filled here). */
EDA_API int eda_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2,
                 int *idx, void *stream);
EDA_API int eda_three_interpolate(const float *points, const int *idx, const float *weight, int B, int C,
                          int m, int n, float *out, void *stream);
EDA_API int eda_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int B,
                               int C, int n, int m, float *grad_points, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EDA_B200_H_ */
