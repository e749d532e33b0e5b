// One-launch all-reduce (sum) of SMALL vectors across the GPUs of one node through NVLink peer memory — the
// cross-rank step of synchronised BatchNorm (the reference converts every BatchNorm to SyncBatchNorm on > 1 GPU,
// main_utils.py:335-338, whose forward all-gathers (mean, invstd, count) and whose backward all-reduces
// (sum_dy, sum_dy_xmu) through NCCL: 2 library collectives per layer and step, 44 on this path).
//
// Here the statistics are 2C numbers per layer and direction; a library collective costs far more in launch and
// protocol latency than the data is worth.  Every rank owns one symmetric buffer (torch symmetric memory: each buffer
// is mapped into every process of the node), laid out as
//     [ seq | err | flags[2][kMaxWorld] | slots[2][world][max_elems] (8-byte elements) ]
// and one single-CTA kernel per reduction does, on every rank at the same point of its stream:
//     1. epoch = ++seq (local word; all ranks run the same sequence, so epochs agree); parity = epoch & 1
//     2. store the local vector into slot [parity][my rank] of EVERY rank's buffer (plain stores over NVLink)
//     3. fence, then one thread per peer writes flags[parity][my rank] = epoch there (st.release.sys)
//     4. one thread per peer spins on the LOCAL flags[parity][peer] until it reads epoch (ld.acquire.sys)
//     5. sum the `world` local slots in rank order — every rank adds the same numbers in the same order, so the
//        result is bit-identical everywhere — and write it over the input.
// Double buffering by parity is enough: a rank can enter epoch e+2 (same parity as e) only after it has seen every
// peer's flag for e+1, which a peer raises after its epoch-e kernel finished reading (stream order).
// A bounded wait (two seconds on the global timer, then sticky) turns a missing peer into an error word instead of
// a hung GPU.
//
// eda_bn_finalize_peer fuses this exchange into the BatchNorm finalisation (sum -> scale / shift / running statistics):
// one launch where the library path needs a collective plus a kernel.
#include "common.cuh"

namespace eda {
namespace {

constexpr int kMaxWorld = 16;
constexpr int kPeerThreads = 256;
constexpr size_t kHeaderBytes = 256;  // seq @0, err @4, flags @64 (2 * 16 * 4 = 128 bytes)

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ double ld_volatile_f64(const double *p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Exchange + ordered sum.  `vals` (n doubles in shared memory) holds the local vector on entry and the sum over the
// ranks on exit (valid for all threads after the function returns).  Returns false on timeout.
__device__ bool peer_exchange_sum(char *const *__restrict__ bufs, int world, int rank, int max_elems, double *vals, int n) {
  __shared__ unsigned s_epoch;
  __shared__ int s_ok;
  char *mine = bufs[rank];
  if (threadIdx.x == 0) {
    unsigned *seq = reinterpret_cast<unsigned *>(mine);
    s_epoch = *seq + 1u;
    *seq = s_epoch;
    s_ok = 1;
  }
  __syncthreads();
  const unsigned epoch = s_epoch;
  const unsigned parity = epoch & 1u;
  const size_t slot_elems = (size_t)max_elems;
  // 2. push
  for (int r = 0; r < world; ++r) {
    double *dst = reinterpret_cast<double *>(bufs[r] + kHeaderBytes) + ((size_t)parity * world + rank) * slot_elems;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = vals[i];
  }
  __threadfence_system();
  __syncthreads();
  // 3. signal, 4. wait
  if (threadIdx.x < world) {
    const int r = threadIdx.x;
    unsigned *flag_there = reinterpret_cast<unsigned *>(bufs[r] + 64) + parity * kMaxWorld + rank;
    st_release_sys(flag_there, epoch);
    const unsigned *flag_here = reinterpret_cast<const unsigned *>(mine + 64) + parity * kMaxWorld + r;
    volatile unsigned *err = reinterpret_cast<volatile unsigned *>(mine) + 1;
    const unsigned long long t0 = global_ns();
    // the error word is sticky: once a peer has failed to arrive, later exchanges do not wait again (a broken run
    // finishes with wrong statistics and a non-zero error word instead of stalling for a second per layer)
    while (ld_acquire_sys(flag_here) != epoch) {
      if (*err != 0u) { s_ok = 0; break; }
      __nanosleep(32);
      if (global_ns() - t0 > 2000000000ull) {  // 2 s: a peer never arrived
        *err = epoch;
        s_ok = 0;
        break;
      }
    }
  }
  __syncthreads();
  // 5. ordered sum
  const double *slots = reinterpret_cast<const double *>(mine + kHeaderBytes) + (size_t)parity * world * slot_elems;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < world; ++r) acc += ld_volatile_f64(slots + (size_t)r * slot_elems + i);
    vals[i] = acc;
  }
  __syncthreads();
  return s_ok != 0;
}

template <typename T>
__global__ void __launch_bounds__(kPeerThreads)
peer_allreduce_kernel(char *const *__restrict__ bufs, int world, int rank, int max_elems, T *__restrict__ data, int n) {
  extern __shared__ double s_vals[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_vals[i] = (double)data[i];
  __syncthreads();
  peer_exchange_sum(bufs, world, rank, max_elems, s_vals, n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) data[i] = (T)s_vals[i];
}

// BatchNorm finalisation with the cross-rank sum of [sum z (C), sum z^2 (C)] fused in (see sa_mlp.cu bn_finalize_kernel
// for the single-rank version: same arithmetic on the summed statistics, count = rows of all ranks).
__global__ void __launch_bounds__(kPeerThreads)
bn_finalize_peer_kernel(char *const *__restrict__ bufs, int world, int rank, int max_elems,
                        const double *__restrict__ stats, double count, const float *__restrict__ gamma,
                        const float *__restrict__ beta, float eps, float momentum, float *__restrict__ running_mean,
                        float *__restrict__ running_var, int update_running, int C, float *__restrict__ scale,
                        float *__restrict__ shift, float *__restrict__ save_mean, float *__restrict__ save_invstd) {
  extern __shared__ double s_vals[];
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_vals[i] = stats[i];
  __syncthreads();
  peer_exchange_sum(bufs, world, rank, max_elems, s_vals, 2 * C);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double mean = s_vals[c] / count;
    double var = s_vals[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double invstd = 1.0 / sqrt(var + (double)eps);
    const float sc = (float)((double)gamma[c] * invstd);
    scale[c] = sc;
    shift[c] = (float)((double)beta[c] - mean * (double)gamma[c] * invstd);
    if (save_mean) save_mean[c] = (float)mean;
    if (save_invstd) save_invstd[c] = (float)invstd;
    if (update_running) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + (double)momentum * mean);
      running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + (double)momentum * unbiased);
    }
  }
}

}  // namespace
}  // namespace eda

extern "C" {

size_t eda_peer_buffer_bytes(int world, int max_elems) {
  if (world < 1 || world > eda::kMaxWorld || max_elems < 1) return 0;
  return eda::kHeaderBytes + (size_t)2 * world * max_elems * sizeof(double);
}

int eda_peer_allreduce(void *const *peer_buffers_dev, int world, int rank, int max_elems, void *data, int n, int is_f64,
                       void *stream) {
  using namespace eda;
  if (!peer_buffers_dev || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || n < 0 || n > max_elems)
    return EDA_ERR_INVALID_ARGUMENT;
  if (n == 0) return EDA_OK;
  if (!data) return EDA_ERR_INVALID_ARGUMENT;
  const size_t smem = (size_t)n * sizeof(double);
  if (smem > 40 * 1024) return EDA_ERR_UNSUPPORTED;
  char *const *bufs = reinterpret_cast<char *const *>(peer_buffers_dev);
  if (is_f64)
    peer_allreduce_kernel<double><<<1, kPeerThreads, smem, as_stream(stream)>>>(bufs, world, rank, max_elems, (double *)data, n);
  else
    peer_allreduce_kernel<float><<<1, kPeerThreads, smem, as_stream(stream)>>>(bufs, world, rank, max_elems, (float *)data, n);
  return check_launch("peer_allreduce_kernel");
}

int eda_bn_finalize_peer(void *const *peer_buffers_dev, int world, int rank, int max_elems, const double *stats,
                         double count, const float *gamma, const float *beta, float eps, float momentum,
                         float *running_mean, float *running_var, int update_running, int C, float *scale, float *shift,
                         float *save_mean, float *save_invstd, void *stream) {
  using namespace eda;
  if (!peer_buffers_dev || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || C < 1 || 2 * C > max_elems)
    return EDA_ERR_INVALID_ARGUMENT;
  if (!stats || !gamma || !beta || !scale || !shift || count <= 0.0) return EDA_ERR_INVALID_ARGUMENT;
  if (update_running && (!running_mean || !running_var)) return EDA_ERR_INVALID_ARGUMENT;
  const size_t smem = (size_t)2 * C * sizeof(double);
  if (smem > 40 * 1024) return EDA_ERR_UNSUPPORTED;
  bn_finalize_peer_kernel<<<1, kPeerThreads, smem, as_stream(stream)>>>(
      reinterpret_cast<char *const *>(peer_buffers_dev), world, rank, max_elems, stats, count, gamma, beta, eps, momentum,
      running_mean, running_var, update_running, C, scale, shift, save_mean, save_invstd);
  return check_launch("bn_finalize_peer_kernel");
}

}  // extern "C"
