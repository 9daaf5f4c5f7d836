// adam.cu -- fused multi-tensor Adam for the 46 parameter tensors of the model.
//
// Reference: torch.optim.Adam(model.parameters(), lr, weight_decay) + optimizer.step() every iteration
// (cad_recognition/train.py:212, :284), i.e. ~100 small launches per step in the eager optimizer.  Here the whole
// update is two launches over a chunk table: a one-thread "tick" that advances the step counter on the device (so the
// update is capturable in the step's CUDA graph) and derives the bias corrections in double precision, then one kernel
// that walks every tensor in 1024-element chunks:
//     g  = grad * grad_scale + weight_decay * p                (L2 penalty folded into the gradient, as torch Adam)
//     m  = beta1 m + (1 - beta1) g ;  v = beta2 v + (1 - beta2) g^2
//     p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// Bound: HBM (16 bytes read + 12 bytes written per element; 45 MB per step for 1.6 M parameters).
#include <cmath>
#include "common.cuh"

namespace yolat {

constexpr int ADAM_CHUNK = 1024;

// state: [0] step (as double), [1] step_size = lr / bc1, [2] 1 / sqrt(bc2)
__global__ void k_adam_tick(double* __restrict__ state, double lr, double beta1, double beta2) {
  const double t = state[0] + 1.0;
  state[0] = t;
  state[1] = lr / (1.0 - pow(beta1, t));
  state[2] = 1.0 / sqrt(1.0 - pow(beta2, t));
}

// Device-resident hyper-parameters (yolat_adam_step_dev): state[3..8] = lr, beta1, beta2, eps, weight_decay, grad_scale.
// A step captured in a CUDA graph reads them at replay time, so a learning-rate schedule (train.py:214 StepLR) or a
// restored checkpoint keeps working after capture: the host only rewrites six doubles in `state`.
__global__ void k_adam_tick_dev(double* __restrict__ state) {
  const double t = state[0] + 1.0;
  state[0] = t;
  state[1] = state[3] / (1.0 - pow(state[4], t));
  state[2] = 1.0 / sqrt(1.0 - pow(state[5], t));
}

// table: per chunk (param*, grad*, m*, v*) as 4 consecutive 64-bit addresses + count
__global__ void __launch_bounds__(256) k_adam_apply(const uint64_t* __restrict__ table, const int32_t* __restrict__ count,
                                                    const double* __restrict__ state, float beta1, float beta2, float omb1,
                                                    float omb2, float eps, float weight_decay, float grad_scale) {
  const uint64_t* e = table + (int64_t)blockIdx.x * 4;
  float* p = reinterpret_cast<float*>(e[0]);
  const float* g = reinterpret_cast<const float*>(e[1]);
  float* m = reinterpret_cast<float*>(e[2]);
  float* v = reinterpret_cast<float*>(e[3]);
  const int n = count[blockIdx.x];
  const float step_size = (float)state[1], inv_sqrt_bc2 = (float)state[2];
  for (int i = threadIdx.x; i < n; i += 256) {
    const float pp = p[i];
    const float gg = fmaf(weight_decay, pp, g[i] * grad_scale);
    const float mm = fmaf(beta1, m[i], omb1 * gg);          // omb = 1 - beta, formed in double on the host as torch does
    const float vv = fmaf(beta2, v[i], omb2 * gg * gg);
    m[i] = mm;
    v[i] = vv;
    p[i] = pp - step_size * (mm / (sqrtf(vv) * inv_sqrt_bc2 + eps));
  }
}

__global__ void __launch_bounds__(256) k_adam_apply_dev(const uint64_t* __restrict__ table, const int32_t* __restrict__ count,
                                                        const double* __restrict__ state) {
  const uint64_t* e = table + (int64_t)blockIdx.x * 4;
  float* p = reinterpret_cast<float*>(e[0]);
  const float* g = reinterpret_cast<const float*>(e[1]);
  float* m = reinterpret_cast<float*>(e[2]);
  float* v = reinterpret_cast<float*>(e[3]);
  const int n = count[blockIdx.x];
  const float step_size = (float)state[1], inv_sqrt_bc2 = (float)state[2];
  const float beta1 = (float)state[4], beta2 = (float)state[5], omb1 = (float)(1.0 - state[4]), omb2 = (float)(1.0 - state[5]);
  const float eps = (float)state[6], weight_decay = (float)state[7], grad_scale = (float)state[8];
  for (int i = threadIdx.x; i < n; i += 256) {
    const float pp = p[i];
    const float gg = fmaf(weight_decay, pp, g[i] * grad_scale);
    const float mm = fmaf(beta1, m[i], omb1 * gg);
    const float vv = fmaf(beta2, v[i], omb2 * gg * gg);
    m[i] = mm;
    v[i] = vv;
    p[i] = pp - step_size * (mm / (sqrtf(vv) * inv_sqrt_bc2 + eps));
  }
}

}  // namespace yolat

using namespace yolat;

extern "C" {

int yolat_adam_chunk(void) { return ADAM_CHUNK; }

int yolat_adam_step(const uint64_t* table, const int32_t* count, int64_t n_chunks, double* state, double lr, double beta1,
                    double beta2, double eps, double weight_decay, double grad_scale, void* stream) {
  if (n_chunks < 0 || !state || (n_chunks > 0 && (!table || !count))) return YOLAT_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  k_adam_tick<<<1, 1, 0, st>>>(state, lr, beta1, beta2);
  YOLAT_CHECK_LAUNCH();
  if (n_chunks > 0) {
    k_adam_apply<<<(unsigned)n_chunks, 256, 0, st>>>(table, count, state, (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps,
                                                     (float)weight_decay, (float)grad_scale);
    YOLAT_CHECK_LAUNCH();
  }
  return YOLAT_OK;
}

// Same update with the hyper-parameters read from state[3..8] on the device (see k_adam_tick_dev): nothing but
// addresses is frozen into a captured graph.
int yolat_adam_step_dev(const uint64_t* table, const int32_t* count, int64_t n_chunks, double* state, void* stream) {
  if (n_chunks < 0 || !state || (n_chunks > 0 && (!table || !count))) return YOLAT_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  k_adam_tick_dev<<<1, 1, 0, st>>>(state);
  YOLAT_CHECK_LAUNCH();
  if (n_chunks > 0) {
    k_adam_apply_dev<<<(unsigned)n_chunks, 256, 0, st>>>(table, count, state);
    YOLAT_CHECK_LAUNCH();
  }
  return YOLAT_OK;
}

}  // extern "C"
