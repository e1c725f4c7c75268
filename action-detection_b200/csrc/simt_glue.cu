// Memory-bound glue of the BNInception graph over NHWC views: layout conversion, max / average
// pooling (Caffe ceil-mode semantics, model_zoo/bninception/layer_factory.py:41-53), 7x7 global
// pooling (bn_inception.yaml:552), ReLU gradient masks, and BN-folding weight packing
// (frozen BatchNorm2d, ssn_models.py:156-174).
#include "common.cuh"

namespace ssnb {
namespace {

constexpr int TPB = 256;
inline unsigned blocks_for(long long n) { return (unsigned)((n + TPB - 1) / TPB); }

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int F, int C, int H, int W, T* __restrict__ dst,
                                    int pitch, int coff, float scale) {
  const long long total = (long long)F * H * W * C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long p = i / C;           // f*H*W + y*W + x
  const long long f = p / ((long long)H * W);
  const long long yx = p % ((long long)H * W);
  dst[p * pitch + coff + c] = from_f<T>(src[(f * C + c) * (long long)H * W + yx] * scale);
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, int F, int C, int H, int W, int pitch, int coff,
                                    float scale, float* __restrict__ dst) {
  const long long total = (long long)F * C * H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long yx = i % ((long long)H * W);
  const int c = (int)((i / ((long long)H * W)) % C);
  const long long f = i / ((long long)H * W * C);
  dst[i] = to_f<T>(src[(f * H * W + yx) * pitch + coff + c]) * scale;
}

// max pooling, ceil_mode: OH = ceil((H + 2p - k)/s) + 1 (last window must start inside the padded input)
template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ src, int H, int W, int C, int spitch, int scoff,
                                   T* __restrict__ dst, int OH, int OW, int dpitch, int dcoff, int F, int k, int stride,
                                   int pad, uint8_t* __restrict__ argmax) {
  const long long total = (long long)F * OH * OW * C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long p = i / C;
  const int ox = (int)(p % OW), oy = (int)((p / OW) % OH);
  const long long f = p / ((long long)OW * OH);
  float best = -INFINITY;
  int bi = 0;
  bool first = true;
  for (int r = 0; r < k; ++r) {
    const int iy = oy * stride + r - pad;
    if (iy < 0 || iy >= H) continue;
    for (int s = 0; s < k; ++s) {
      const int ix = ox * stride + s - pad;
      if (ix < 0 || ix >= W) continue;
      const float v = to_f<T>(src[((f * H + iy) * W + ix) * spitch + scoff + c]);
      if (first || v > best || v != v) { best = v; bi = r * k + s; first = false; }   // first max wins (ATen)
    }
  }
  dst[p * dpitch + dcoff + c] = from_f<T>(best);
  argmax[i] = (uint8_t)bi;
}

template <typename T>
__global__ void maxpool_bwd_kernel(T* __restrict__ dsrc, int H, int W, int C, int spitch, int scoff,
                                   const T* __restrict__ ddst, int OH, int OW, int dpitch, int dcoff, int F, int k,
                                   int stride, int pad, const uint8_t* __restrict__ argmax, int accumulate) {
  const long long total = (long long)F * H * W * C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long p = i / C;
  const int ix = (int)(p % W), iy = (int)((p / W) % H);
  const long long f = p / ((long long)W * H);
  float g = 0.f;
  for (int r = 0; r < k; ++r) {
    const int ty = iy + pad - r;
    if (ty < 0 || ty % stride) continue;
    const int oy = ty / stride;
    if (oy >= OH) continue;
    for (int s = 0; s < k; ++s) {
      const int tx = ix + pad - s;
      if (tx < 0 || tx % stride) continue;
      const int ox = tx / stride;
      if (ox >= OW) continue;
      const long long op = (f * OH + oy) * OW + ox;
      if (argmax[op * C + c] == (uint8_t)(r * k + s)) g += to_f<T>(ddst[op * dpitch + dcoff + c]);
    }
  }
  T* q = dsrc + p * spitch + scoff + c;
  if (accumulate) g += to_f<T>(*q);
  *q = from_f<T>(g);
}

// 3x3 stride-1 pad-1 average, count_include_pad (always /9); its own adjoint
template <typename T>
__global__ void avgpool3_kernel(const T* __restrict__ src, int H, int W, int C, int spitch, int scoff,
                                T* __restrict__ dst, int dpitch, int dcoff, int F, int accumulate) {
  const long long total = (long long)F * H * W * C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long p = i / C;
  const int x = (int)(p % W), y = (int)((p / W) % H);
  const long long f = p / ((long long)W * H);
  float s = 0.f;
  for (int r = -1; r <= 1; ++r) {
    const int yy = y + r;
    if (yy < 0 || yy >= H) continue;
    for (int q = -1; q <= 1; ++q) {
      const int xx = x + q;
      if (xx < 0 || xx >= W) continue;
      s += to_f<T>(src[((f * H + yy) * W + xx) * spitch + scoff + c]);
    }
  }
  s = s / 9.0f;
  T* o = dst + p * dpitch + dcoff + c;
  if (accumulate) s += to_f<T>(*o);
  *o = from_f<T>(s);
}

template <typename T>
__global__ void gpool_fwd_kernel(const T* __restrict__ src, int HW, int C, int pitch, int coff, int F,
                                 float* __restrict__ feat) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)F * C) return;
  const int c = (int)(i % C);
  const long long f = i / C;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += to_f<T>(src[(f * HW + p) * pitch + coff + c]);
  feat[i] = s / (float)HW;
}

template <typename T>
__global__ void gpool_bwd_kernel(const float* __restrict__ dfeat, float scale, int HW, int C, int pitch, int coff, int F,
                                 T* __restrict__ ddst, const T* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)F * HW * C) return;
  const int c = (int)(i % C);
  const long long p = i / C;
  const long long f = p / HW;
  float g = dfeat[f * C + c] / (float)HW * scale;
  if (y && !(to_f<T>(y[p * pitch + coff + c]) > 0.f)) g = 0.f;     // fused ReLU gradient mask (same view geometry)
  ddst[p * pitch + coff + c] = from_f<T>(g);
}

template <typename T>
__global__ void relu_mask_kernel(T* __restrict__ dy, int dpitch, int dcoff, const T* __restrict__ y, int ypitch,
                                 int ycoff, long long pixels, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels * C) return;
  const int c = (int)(i % C);
  const long long p = i / C;
  if (!(to_f<T>(y[p * ypitch + ycoff + c]) > 0.f)) dy[p * dpitch + dcoff + c] = from_f<T>(0.f);
}

template <typename T>
__global__ void fill_zero_kernel(T* __restrict__ v, int pitch, int coff, long long pixels, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels * C) return;
  v[(i / C) * pitch + coff + (int)(i % C)] = from_f<T>(0.f);
}

// fold BN: s = gamma / sqrt(var + eps); W' = W*s ; b' = (b - mean)*s + beta
template <typename T>
__global__ void pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ mean, const float* __restrict__ var, int Cout, int Cin, int k,
                                 T* __restrict__ wf, T* __restrict__ wd, float* __restrict__ bias_f,
                                 float* __restrict__ scale, float* __restrict__ absmax) {
  const int taps = k * k;
  const long long total = (long long)Cout * Cin * taps;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Cout) {
    const float s = gamma[i] / sqrtf(var[i] + 1e-5f);
    scale[i] = s;
    bias_f[i] = (b[i] - mean[i]) * s + beta[i];
  }
  float av = 0.f;
  if (i < total) {
    const int tap = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const int co = (int)(i / ((long long)taps * Cin));
    const float s = gamma[co] / sqrtf(var[co] + 1e-5f);
    const float fv = w[i] * s;
    const T v = from_f<T>(fv);
    wf[((long long)tap * Cin + ci) * Cout + co] = v;
    wd[((long long)tap * Cout + co) * Cin + ci] = v;
    av = fabsf(fv);
  }
  if (absmax) {          // SSNB_EXACT_TC: largest folded weight of the layer (non-negative floats order like their bit patterns)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) av = fmaxf(av, __shfl_xor_sync(0xffffffffu, av, o));
    if (threadIdx.x % 32 == 0 && av > 0.f) atomicMax(reinterpret_cast<int*>(absmax), __float_as_int(av));
  }
}

// all layers of the network in a few launches (the 69 per-layer launches cost ~1.1 ms per training step, mostly launch
// latency: the weights are re-packed after every optimizer step)
template <typename T>
__global__ void pack_all_kernel(const __grid_constant__ PackTable t) {
  int ei = 0;
  while (ei + 1 < t.n && (int)blockIdx.x >= t.e[ei + 1].block0) ++ei;       // <= 32 entries, uniform per block
  const PackEntry& q = t.e[ei];
  const int taps = q.k * q.k;
  const long long total = (long long)q.cout * q.cin * taps;
  float av = 0.f;
  __shared__ float wmax[TPB / 32];
  if (taps <= PACK_TILE_TAPS) {
    // tiled path: a CTA owns PACK_TILE output x PACK_TILE input channels x taps.  The source rows ([ci][tap] runs of one output
    // channel) are read contiguously into shared memory, wf is then written with the output channel and wd with the input channel
    // as the lane index, so all three streams are coalesced (the element-wise path scatters both stores: 0.32 ms per step)
    __shared__ float tile[PACK_TILE][PACK_TILE * PACK_TILE_TAPS + 1];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int ci_tiles = (q.cin + PACK_TILE - 1) / PACK_TILE;
    const int b = (int)blockIdx.x - q.block0;
    const int co0 = (b / ci_tiles) * PACK_TILE, ci0 = (b % ci_tiles) * PACK_TILE;
    const int nco = min(PACK_TILE, q.cout - co0), nci = min(PACK_TILE, q.cin - ci0);
    const int row = nci * taps;
    if (ci0 == 0 && (int)threadIdx.x < nco) {
      const int co = co0 + threadIdx.x;
      const float s = q.nofold ? 1.0f : q.gamma[co] / sqrtf(q.var[co] + 1e-5f);
      q.scale[co] = s;
      const float bf = q.nofold ? q.b[co] : (q.b[co] - q.mean[co]) * s + q.beta[co];
      q.bias[co] = bf;
      if (q.bias_b) q.bias_b[co] = bf;
    }
    for (int r = warp; r < nco; r += TPB / 32) {
      const int co = co0 + r;
      const float s = q.nofold ? 1.0f : q.gamma[co] / sqrtf(q.var[co] + 1e-5f);
      const float* src = q.w + ((long long)co * q.cin + ci0) * taps;
      for (int e = lane; e < row; e += 32) { const float fv = src[e] * s; tile[r][e] = fv; av = fmaxf(av, fabsf(fv)); }
    }
    __syncthreads();
    T* wf = reinterpret_cast<T*>(q.wf);
    T* wd = reinterpret_cast<T*>(q.wd);
    for (int e = warp; e < row; e += TPB / 32) {                 // e = local ci * taps + tap; lanes = output channels
      const int cil = e / taps, tap = e - cil * taps;
      if (lane < nco) wf[((long long)tap * q.cin + ci0 + cil) * q.cout + co0 + lane] = from_f<T>(tile[lane][e]);
    }
    for (int e = warp; e < nco * taps; e += TPB / 32) {          // e = local co * taps + tap; lanes = input channels
      const int r = e / taps, tap = e - r * taps;
      if (lane < nci) wd[((long long)tap * q.cout + co0 + r) * q.cin + ci0 + lane] = from_f<T>(tile[r][lane * taps + tap]);
    }
  } else {
  const long long base = (long long)((int)blockIdx.x - q.block0) * (PACK_PER_THREAD * TPB) + threadIdx.x;
#pragma unroll
  for (int j = 0; j < PACK_PER_THREAD; ++j) {
    const long long i = base + j * TPB;
    if (i < q.cout) {
      const float s = q.nofold ? 1.0f : q.gamma[i] / sqrtf(q.var[i] + 1e-5f);
      q.scale[i] = s;
      const float bf = q.nofold ? q.b[i] : (q.b[i] - q.mean[i]) * s + q.beta[i];
      q.bias[i] = bf;
      if (q.bias_b) q.bias_b[i] = bf;
    }
    if (i < total) {
      const int tap = (int)(i % taps);
      const int ci = (int)((i / taps) % q.cin);
      const int co = (int)(i / ((long long)taps * q.cin));
      const float s = q.nofold ? 1.0f : q.gamma[co] / sqrtf(q.var[co] + 1e-5f);
      const float fv = q.w[i] * s;
      const T v = from_f<T>(fv);
      reinterpret_cast<T*>(q.wf)[((long long)tap * q.cin + ci) * q.cout + co] = v;
      reinterpret_cast<T*>(q.wd)[((long long)tap * q.cout + co) * q.cin + ci] = v;
      av = fmaxf(av, fabsf(fv));
    }
  }
  }
  if (q.absmax) {
    // one atomic per CTA instead of one per warp (they serialise in L2 on the 69 per-layer addresses)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) av = fmaxf(av, __shfl_xor_sync(0xffffffffu, av, o));
    if (threadIdx.x % 32 == 0) wmax[threadIdx.x / 32] = av;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w = 1; w < TPB / 32; ++w) av = fmaxf(av, wmax[w]);
      if (av > 0.f) atomicMax(reinterpret_cast<int*>(q.absmax), __float_as_int(av));
    }
  }
}

}  // namespace

template <typename T> int launch_pack_all(const PackTable& t, int total_blocks, cudaStream_t s) {
  if (t.n <= 0) return 0;
  pack_all_kernel<T><<<(unsigned)total_blocks, TPB, 0, s>>>(t);
  SSNB_LAUNCH_CHECK("pack_all_kernel");
  return 0;
}
template int launch_pack_all<float>(const PackTable&, int, cudaStream_t);
template int launch_pack_all<__half>(const PackTable&, int, cudaStream_t);

namespace {
}  // namespace

#define V(T, v) reinterpret_cast<T*>((v).base)

template <typename T> int launch_nchw_to_nhwc(const float* src, int F, int C, int H, int W, View dst, float scale, cudaStream_t s) {
  const long long n = (long long)F * C * H * W;
  nchw_to_nhwc_kernel<T><<<blocks_for(n), TPB, 0, s>>>(src, F, C, H, W, V(T, dst), dst.pitch, dst.coff, scale);
  SSNB_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return 0;
}
template <typename T> int launch_nhwc_to_nchw(View src, int F, float scale, float* dst, cudaStream_t s) {
  const long long n = (long long)F * src.C * src.H * src.W;
  nhwc_to_nchw_kernel<T><<<blocks_for(n), TPB, 0, s>>>(V(const T, src), F, src.C, src.H, src.W, src.pitch, src.coff, scale, dst);
  SSNB_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return 0;
}
template <typename T>
int launch_maxpool_fwd(View src, View dst, int F, int k, int stride, int pad, uint8_t* argmax, cudaStream_t s) {
  const long long n = (long long)F * dst.H * dst.W * src.C;
  maxpool_fwd_kernel<T><<<blocks_for(n), TPB, 0, s>>>(V(const T, src), src.H, src.W, src.C, src.pitch, src.coff, V(T, dst),
                                                     dst.H, dst.W, dst.pitch, dst.coff, F, k, stride, pad, argmax);
  SSNB_LAUNCH_CHECK("maxpool_fwd_kernel");
  return 0;
}
template <typename T>
int launch_maxpool_bwd(View dsrc, View ddst, int F, int k, int stride, int pad, const uint8_t* argmax, int accumulate,
                       cudaStream_t s) {
  const long long n = (long long)F * dsrc.H * dsrc.W * dsrc.C;
  maxpool_bwd_kernel<T><<<blocks_for(n), TPB, 0, s>>>(V(T, dsrc), dsrc.H, dsrc.W, dsrc.C, dsrc.pitch, dsrc.coff,
                                                     V(const T, ddst), ddst.H, ddst.W, ddst.pitch, ddst.coff, F, k, stride,
                                                     pad, argmax, accumulate);
  SSNB_LAUNCH_CHECK("maxpool_bwd_kernel");
  return 0;
}
template <typename T> int launch_avgpool3_fwd(View src, View dst, int F, int accumulate, cudaStream_t s) {
  const long long n = (long long)F * src.H * src.W * src.C;
  avgpool3_kernel<T><<<blocks_for(n), TPB, 0, s>>>(V(const T, src), src.H, src.W, src.C, src.pitch, src.coff, V(T, dst),
                                                  dst.pitch, dst.coff, F, accumulate);
  SSNB_LAUNCH_CHECK("avgpool3_kernel");
  return 0;
}
template <typename T> int launch_gpool_fwd(View src, int F, float* feat, cudaStream_t s) {
  gpool_fwd_kernel<T><<<blocks_for((long long)F * src.C), TPB, 0, s>>>(V(const T, src), src.H * src.W, src.C, src.pitch,
                                                                      src.coff, F, feat);
  SSNB_LAUNCH_CHECK("gpool_fwd_kernel");
  return 0;
}
template <typename T> int launch_gpool_bwd(const float* dfeat, float scale, View ddst, int F, const void* y, cudaStream_t s) {
  const long long n = (long long)F * ddst.H * ddst.W * ddst.C;
  gpool_bwd_kernel<T><<<blocks_for(n), TPB, 0, s>>>(dfeat, scale, ddst.H * ddst.W, ddst.C, ddst.pitch, ddst.coff, F, V(T, ddst), reinterpret_cast<const T*>(y));
  SSNB_LAUNCH_CHECK("gpool_bwd_kernel");
  return 0;
}
template <typename T> int launch_relu_mask(View dy, View y, int F, cudaStream_t s) {
  const long long px = (long long)F * dy.H * dy.W;
  relu_mask_kernel<T><<<blocks_for(px * dy.C), TPB, 0, s>>>(V(T, dy), dy.pitch, dy.coff, V(const T, y), y.pitch, y.coff, px, dy.C);
  SSNB_LAUNCH_CHECK("relu_mask_kernel");
  return 0;
}
template <typename T> int launch_fill_zero(View v, int F, cudaStream_t s) {
  const long long px = (long long)F * v.H * v.W;
  fill_zero_kernel<T><<<blocks_for(px * v.C), TPB, 0, s>>>(V(T, v), v.pitch, v.coff, px, v.C);
  SSNB_LAUNCH_CHECK("fill_zero_kernel");
  return 0;
}
template <typename T>
int launch_pack_conv(const float* w, const float* b, const float* gamma, const float* beta, const float* mean,
                     const float* var, int Cout, int Cin, int k, T* wf, T* wd, float* bias_f, float* scale,
                     cudaStream_t s, float* absmax) {
  const long long n = (long long)Cout * Cin * k * k;
  pack_conv_kernel<T><<<blocks_for(n > Cout ? n : Cout), TPB, 0, s>>>(w, b, gamma, beta, mean, var, Cout, Cin, k, wf, wd, bias_f, scale, absmax);
  SSNB_LAUNCH_CHECK("pack_conv_kernel");
  return 0;
}

#define INST(T)                                                                                              \
  template int launch_nchw_to_nhwc<T>(const float*, int, int, int, int, View, float, cudaStream_t);                 \
  template int launch_nhwc_to_nchw<T>(View, int, float, float*, cudaStream_t);                               \
  template int launch_maxpool_fwd<T>(View, View, int, int, int, int, uint8_t*, cudaStream_t);                \
  template int launch_maxpool_bwd<T>(View, View, int, int, int, int, const uint8_t*, int, cudaStream_t);     \
  template int launch_avgpool3_fwd<T>(View, View, int, int, cudaStream_t);                                   \
  template int launch_gpool_fwd<T>(View, int, float*, cudaStream_t);                                         \
  template int launch_gpool_bwd<T>(const float*, float, View, int, const void*, cudaStream_t);                            \
  template int launch_relu_mask<T>(View, View, int, cudaStream_t);                                           \
  template int launch_fill_zero<T>(View, int, cudaStream_t);                                                 \
  template int launch_pack_conv<T>(const float*, const float*, const float*, const float*, const float*,     \
                                   const float*, int, int, int, T*, T*, float*, float*, cudaStream_t, float*);
INST(float)
INST(__half)

}  // namespace ssnb
