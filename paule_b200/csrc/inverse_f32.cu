// Stencil kernels of InverseModelMelTimeSmoothResidual (/root/reference/paule/models.py:177-247).
// Forward only, once per word (initial cp from the target mel, paule/paule.py:551-557); batch-first.
#include "common.cuh"

namespace paule {

// MelChannelConv1D + residual (models.py:152-169, :224-228):
//   y[b,t,c] = x[b,t,c] + bias[c] + sum_{dm<3, dt<5} w[c][dm][dt] * x[b, t+dt-2, c+dm-1]   (zero padded)
__global__ void __launch_bounds__(256)
melconv_res_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ y, int64_t B, int64_t Tm, int64_t Cm) {
  const int64_t n = B * Tm * Cm;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = e % Cm, t = (e / Cm) % Tm, b = e / (Cm * Tm);
    const float* xb = x + b * Tm * Cm;
    float acc = 0.f;
#pragma unroll
    for (int dm = 0; dm < 3; ++dm) {
      const int64_t cc = c + dm - 1;
      if (cc < 0 || cc >= Cm) continue;
#pragma unroll
      for (int dt = 0; dt < 5; ++dt) {
        const int64_t tt = t + dt - 2;
        if (tt < 0 || tt >= Tm) continue;
        acc = fmaf(__ldg(w + (c * 3 + dm) * 5 + dt), __ldg(xb + tt * Cm + cc), acc);
      }
    }
    y[e] = (acc + __ldg(bias + c)) + __ldg(x + e);
  }
}

// add_vel_and_acc_info (models.py:47-61)
__global__ void __launch_bounds__(256)
vel_acc_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t B, int64_t Tm, int64_t Cm) {
  const int64_t n = B * Tm * Cm;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = e % Cm, t = (e / Cm) % Tm, b = e / (Cm * Tm);
    const float* xb = x + b * Tm * Cm;
    const float x0 = xb[t * Cm + c];
    const float vel = (t + 1 < Tm) ? xb[(t + 1) * Cm + c] - x0 : 0.f;
    float acc = 0.f;
    if (t >= 1 && t + 1 < Tm) acc = (xb[(t + 1) * Cm + c] - x0) - (x0 - xb[(t - 1) * Cm + c]);
    float* yo = y + (b * Tm + t) * 3 * Cm;
    yo[c] = x0;
    yo[Cm + c] = vel;
    yo[2 * Cm + c] = acc;
  }
}

// double_sequence (models.py:63-81): y[2k] = x[k]; y[2k+1] = (x[k]+x[k+1])/2; y[2Tm-1] = x[Tm-1]
__global__ void __launch_bounds__(256)
double_sequence_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t B, int64_t Tm, int64_t C) {
  const int64_t n = B * 2 * Tm * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = e % C, t2 = (e / C) % (2 * Tm), b = e / (C * 2 * Tm);
    const int64_t k = t2 / 2;
    const float* xb = x + b * Tm * C;
    float v = xb[k * C + c];
    if ((t2 & 1) && k + 1 < Tm) v = (v + xb[(k + 1) * C + c]) / 2.0f;
    y[e] = v;
  }
}

// depthwise 5-tap conv over time, zero padded:  y = conv(x; w[c][5], b[c]) (+ resid)
__global__ void __launch_bounds__(256)
dwconv5_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ resid, float* __restrict__ y, int64_t B, int64_t T, int64_t C) {
  const int64_t n = B * T * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = e % C, t = (e / C) % T, b = e / (C * T);
    const float* xb = x + b * T * C;
    float acc = 0.f;
#pragma unroll
    for (int dt = 0; dt < 5; ++dt) {
      const int64_t tt = t + dt - 2;
      if (tt >= 0 && tt < T) acc = fmaf(__ldg(w + c * 5 + dt), xb[tt * C + c], acc);
    }
    acc += __ldg(bias + c);
    if (resid) acc += resid[e];
    y[e] = acc;
  }
}

// resid_weighting (models.py:217-219, :241-244): per channel 5-tap on the smoothed + 5-tap on the raw signal
__global__ void __launch_bounds__(256)
mix5_kernel(const float* __restrict__ smooth, const float* __restrict__ raw, const float* __restrict__ w,
            const float* __restrict__ bias, float* __restrict__ y, int64_t B, int64_t T, int64_t C) {
  const int64_t n = B * T * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = e % C, t = (e / C) % T, b = e / (C * T);
    const float* sb = smooth + b * T * C;
    const float* rb = raw + b * T * C;
    float acc = 0.f;
#pragma unroll
    for (int dt = 0; dt < 5; ++dt) {
      const int64_t tt = t + dt - 2;
      if (tt >= 0 && tt < T) {
        acc = fmaf(__ldg(w + (c * 2 + 0) * 5 + dt), sb[tt * C + c], acc);
        acc = fmaf(__ldg(w + (c * 2 + 1) * 5 + dt), rb[tt * C + c], acc);
      }
    }
    y[e] = acc + __ldg(bias + c);
  }
}

static inline unsigned grid_for(int64_t n) {
  int64_t blocks = ceil_div(n, (int64_t)256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace paule

using namespace paule;

extern "C" int paule_melconv_res_f32(const float* x, const float* w, const float* bias, float* y, int64_t B,
                                     int64_t Tm, int64_t Cm, paule_stream_t stream) {
  PAULE_REQUIRE(x && w && bias && y && x != y && B > 0 && Tm > 0 && Cm > 0);
  melconv_res_kernel<<<grid_for(B * Tm * Cm), 256, 0, as_stream(stream)>>>(x, w, bias, y, B, Tm, Cm);
  PAULE_LAUNCH_CHECK("melconv_res_kernel");
  return PAULE_OK;
}

extern "C" int paule_vel_acc_f32(const float* x, float* y, int64_t B, int64_t Tm, int64_t Cm, paule_stream_t stream) {
  PAULE_REQUIRE(x && y && B > 0 && Tm > 0 && Cm > 0);
  vel_acc_kernel<<<grid_for(B * Tm * Cm), 256, 0, as_stream(stream)>>>(x, y, B, Tm, Cm);
  PAULE_LAUNCH_CHECK("vel_acc_kernel");
  return PAULE_OK;
}

extern "C" int paule_upsample_smooth_f32(const float* x, const float* res_w, const float* res_b, int n_blocks,
                                         const float* mix_w, const float* mix_b, float* y, float* scratch, int64_t B,
                                         int64_t Tm, int64_t C, paule_stream_t stream) {
  PAULE_REQUIRE(x && y && scratch && B > 0 && Tm > 0 && C > 0 && n_blocks >= 0);
  PAULE_REQUIRE(n_blocks == 0 || (res_w && res_b && mix_w && mix_b));
  const int64_t T2 = 2 * Tm, n = B * T2 * C;
  cudaStream_t s = as_stream(stream);
  const unsigned g = grid_for(n);
  if (n_blocks == 0) {
    double_sequence_kernel<<<g, 256, 0, s>>>(x, y, B, Tm, C);
    PAULE_LAUNCH_CHECK("double_sequence_kernel");
    return PAULE_OK;
  }
  // scratch = raw | tmp | alt (3n floats).  Block outputs alternate between y and alt such that the LAST block
  // lands in alt, so the final mix can write y without aliasing its inputs.
  float* raw = scratch;
  float* tmp = scratch + n;
  float* alt = scratch + 2 * n;
  double_sequence_kernel<<<g, 256, 0, s>>>(x, raw, B, Tm, C);
  const float* cur = raw;
  for (int k = 0; k < n_blocks; ++k) {
    const float* w1 = res_w + ((int64_t)k * 2 + 0) * C * 5;
    const float* w2 = res_w + ((int64_t)k * 2 + 1) * C * 5;
    const float* b1 = res_b + ((int64_t)k * 2 + 0) * C;
    const float* b2 = res_b + ((int64_t)k * 2 + 1) * C;
    float* out = ((n_blocks - k) & 1) ? alt : y;
    dwconv5_kernel<<<g, 256, 0, s>>>(cur, w1, b1, nullptr, tmp, B, T2, C);   // band_conv1d_1
    dwconv5_kernel<<<g, 256, 0, s>>>(tmp, w2, b2, cur, out, B, T2, C);       // band_conv1d_2 + residual
    cur = out;
  }
  mix5_kernel<<<g, 256, 0, s>>>(cur, raw, mix_w, mix_b, y, B, T2, C);
  PAULE_LAUNCH_CHECK("upsample_smooth kernels");
  return PAULE_OK;
}
