// SSNB_EXACT_TC glue: error-compensated fp16 operand planes of fp32 tensors.
//
// The reference computes every convolution in fp32 (model_zoo/bninception/layer_factory.py:25-39, no AMP in
// ssn_train.py:81).  tcgen05 has no fp32 MMA; an fp32 value x is instead carried as TWO fp16 numbers
//     hi = fp16(x),  lo = fp16(x - float(hi))          (hi + lo == x to ~2^-22 relative, fp16 range permitting)
// and a product a*b is evaluated as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi on the tensor cores with fp32 accumulation
// (umma_conv_v2.cu / umma_conv.cu / umma_wgrad.cu, nseg = 3).  The kernels here produce those planes for tensors that
// do not come out of a convolution epilogue (pool outputs, the network input, masked output gradients, weights).
#include "common.cuh"

namespace ssnb {
namespace {

constexpr int TPB = 256;
constexpr float HALF_MAX = 65504.f;

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// one thread = 8 channels of one pixel: 2 x 16-byte fp32 loads -> 16 bytes of hi + 16 bytes of lo
__global__ void split_view_kernel(const float* __restrict__ src, int spitch, int scoff, long long pixels, int C, float scale,
                                  __half* __restrict__ hi, int hpitch, int hcoff, long long lo_off, int* __restrict__ flag) {
  const int G = C / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels * G) return;
  const int g = (int)(i % G);
  const long long p = i / G;
  const float4* s4 = reinterpret_cast<const float4*>(src + p * spitch + scoff + g * 8);
  float4 a = __ldg(s4), b = __ldg(s4 + 1);
  a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale; b.x *= scale; b.y *= scale; b.z *= scale; b.w *= scale;
  if (flag) {
    const float m = fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                          fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
    if (!(m <= HALF_MAX)) *flag = 1;                   // overflow or NaN under this loss scale
  }
  uint4 h, l;
  split2(a.x, a.y, h.x, l.x); split2(a.z, a.w, h.y, l.y); split2(b.x, b.y, h.z, l.z); split2(b.z, b.w, h.w, l.w);
  __half* hp = hi + p * hpitch + hcoff + g * 8;
  *reinterpret_cast<uint4*>(hp) = h;
  *reinterpret_cast<uint4*>(reinterpret_cast<char*>(hp) + lo_off) = l;
}

__global__ void split_flat_kernel(const float* __restrict__ src, long long n, __half* __restrict__ hi, __half* __restrict__ lo,
                                  const float* __restrict__ absmax, float* __restrict__ inv_scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // power-of-two scale (exact in fp): the largest weight lands in [4096, 8192), so hi's ulp is >= 4 and lo >= 2^-... stays
  // a NORMAL fp16 number down to weights 2^-13 below the maximum.  Unscaled, a BN-folded weight of ~1e-3 (conv1: inputs
  // are +-128, the fold divides by sigma ~ 100) has a subnormal lo with 3 significant bits: measured 1.4e-5 error.
  float sc = 1.0f;
  if (absmax) {
    int e = 0;
    const float m = *absmax;
    if (m > 0.f && m < INFINITY) { frexpf(m, &e); sc = ldexpf(1.0f, 13 - e); }
    if (i == 0 && inv_scale) *inv_scale = 1.0f / sc;
  }
  if (i >= n) return;
  const float v = src[i] * sc;
  const __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}

__global__ void split_all_kernel(const __grid_constant__ SplitTable t) {
  int ei = 0;
  while (ei + 1 < t.n && (int)blockIdx.x >= t.e[ei + 1].block0) ++ei;
  const SplitEntry& q = t.e[ei];
  const long long i = (long long)((int)blockIdx.x - q.block0) * blockDim.x + threadIdx.x;
  float sc = 1.0f;
  int e = 0;
  const float m = *q.absmax;
  if (m > 0.f && m < INFINITY) { frexpf(m, &e); sc = ldexpf(1.0f, 13 - e); }
  if (i == 0) *q.inv_scale = 1.0f / sc;
  if (i >= q.n) return;
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const float v = (l ? q.wd : q.wf)[i] * sc;
    __half* hi = l ? q.wd16 : q.wf16;
    const __half h = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(h));
    hi[i] = h;
    *reinterpret_cast<__half*>(reinterpret_cast<char*>(hi + i) + q.plane_bytes) = lo;
    __half* hb = l ? q.wd16_b : q.wf16_b;
    if (hb) {                                   // 1x1 member of a fused sibling block: wd[co][ci] rows stacked, wf[ci][co] as a column block
      const long long j = l ? i : (i / q.cout) * (long long)q.b_pitch + (i % q.cout);
      hb[j] = h;
      *reinterpret_cast<__half*>(reinterpret_cast<char*>(hb + j) + q.b_plane_bytes) = lo;
    }
  }
}

// diagnostic read-back: (hi + lo) * scale as NCHW fp32 (what the consuming tensor-core kernels see)
__global__ void planes_to_nchw_kernel(const __half* __restrict__ hi, long long lo_off, int F, int C, int H, int W, int pitch, int coff,
                                      float scale, float* __restrict__ dst) {
  const long long total = (long long)F * C * H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long yx = i % ((long long)H * W);
  const int c = (int)((i / ((long long)H * W)) % C);
  const long long f = i / ((long long)H * W * C);
  const __half* hp = hi + (f * H * W + yx) * pitch + coff + c;
  const __half* lp = reinterpret_cast<const __half*>(reinterpret_cast<const char*>(hp) + lo_off);
  dst[i] = (__half2float(*hp) + __half2float(*lp)) * scale;
}

// NCHW fp32 frames -> conv1's packed space-to-depth operand planes (layout of s2d_glue.cu: channel = ds*Cs + (a*2+b)*Cin + c
// holds x[f, 2i+a, 2(j+ds-2)+b, c]); one thread per (pixel, ds block)
// CIN > 0: compile-time channel count (RGB 3 / Flow 10): the channel loop unrolls and the staging arrays stay in registers
template <int CS, int CIN>
__global__ void nchw_to_s2d_split_kernel(const float* __restrict__ src, int F, int Cin_rt, int H, int W, __half* __restrict__ dst,
                                         long long lo_off) {
  const int Cin = CIN ? CIN : Cin_rt;
  const int H2 = H / 2, W2 = W / 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)F * H2 * W2 * 4) return;
  const unsigned iu = (unsigned)i;
  const int ds = (int)(iu & 3u);
  const unsigned pu = iu >> 2;
  const int x2 = (int)(pu % (unsigned)W2), y2 = (int)((pu / (unsigned)W2) % (unsigned)H2);
  const long long p = pu;
  const long long f = pu / (unsigned)(W2 * H2);
  __align__(16) __half vh[CS];
  __align__(16) __half vl[CS];
#pragma unroll
  for (int c = 0; c < CS; ++c) { vh[c] = __float2half_rn(0.f); vl[c] = __float2half_rn(0.f); }
  const int xs = x2 + ds - 2;
  if (xs >= 0 && xs < W2) {
#pragma unroll
    for (int c = 0; c < Cin; ++c) {
      const float* pl = src + ((f * Cin + c) * H + 2 * y2) * (long long)W + 2 * xs;
      const float2 r0 = __ldg(reinterpret_cast<const float2*>(pl));
      const float2 r1 = __ldg(reinterpret_cast<const float2*>(pl + W));
      const float q[4] = {r0.x, r0.y, r1.x, r1.y};
#pragma unroll
      for (int ab = 0; ab < 4; ++ab) {
        const __half h = __float2half_rn(q[ab]);
        vh[ab * Cin + c] = h;
        vl[ab * Cin + c] = __float2half_rn(q[ab] - __half2float(h));
      }
    }
  }
  __half* o = dst + p * (4 * CS) + ds * CS;
  uint4* oh = reinterpret_cast<uint4*>(o);
  uint4* ol = reinterpret_cast<uint4*>(reinterpret_cast<char*>(o) + lo_off);
#pragma unroll
  for (int q = 0; q < CS / 8; ++q) { oh[q] = reinterpret_cast<const uint4*>(vh)[q]; ol[q] = reinterpret_cast<const uint4*>(vl)[q]; }
}

// fp32 NHWC view -> the same packed space-to-depth planes (per-layer tests: the input was written as a named value)
__global__ void nhwc_to_s2d_split_kernel(const float* __restrict__ src, int F, int H, int W, int Cin, int spitch, int scoff,
                                         __half* __restrict__ dst, long long lo_off, int Cs) {
  const int H2 = H / 2, W2 = W / 2, Ck = 4 * Cs;
  const long long total = (long long)F * H2 * W2 * Ck;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = (int)(i % Ck);
  const long long p = i / Ck;
  const int x2 = (int)(p % W2), y2 = (int)((p / W2) % H2);
  const long long f = p / ((long long)W2 * H2);
  const int ds = ch / Cs, q = ch % Cs;
  float v = 0.f;
  const int xs = x2 + ds - 2;
  if (q < 4 * Cin && xs >= 0 && xs < W2) {
    const int ab = q / Cin, c = q % Cin;
    v = src[((f * H + 2 * y2 + ab / 2) * W + 2 * xs + ab % 2) * spitch + scoff + c];
  }
  const __half h = __float2half_rn(v);
  dst[i] = h;
  *reinterpret_cast<__half*>(reinterpret_cast<char*>(dst + i) + lo_off) = __float2half_rn(v - __half2float(h));
}

}  // namespace

int launch_nhwc_to_s2d_split(View src, int F, __half* dst_hi, long long lo_off, int Cs, cudaStream_t s) {
  const long long n = (long long)F * (src.H / 2) * (src.W / 2) * 4 * Cs;
  nhwc_to_s2d_split_kernel<<<(unsigned)((n + TPB - 1) / TPB), TPB, 0, s>>>((const float*)src.base, F, src.H, src.W, src.C, src.pitch, src.coff,
                                                                         dst_hi, lo_off, Cs);
  SSNB_LAUNCH_CHECK("nhwc_to_s2d_split_kernel");
  return 0;
}

int launch_split_view(View src, int F, float scale, View planes, int* flag, cudaStream_t s) {
  if (src.C % 8 || src.pitch % 4 || src.coff % 4 || planes.pitch % 8 || planes.coff % 8 || planes.C != src.C || !planes.lo_off) {
    set_thread_error("split_view: channel counts / offsets must be multiples of 8 and the planes view must carry a LO plane"); return 1; }
  const long long px = (long long)F * src.H * src.W;
  const long long n = px * (src.C / 8);
  split_view_kernel<<<(unsigned)((n + TPB - 1) / TPB), TPB, 0, s>>>((const float*)src.base, src.pitch, src.coff, px, src.C, scale,
                                                                  (__half*)planes.base, planes.pitch, planes.coff, planes.lo_off, flag);
  SSNB_LAUNCH_CHECK("split_view_kernel");
  return 0;
}
int launch_split_flat(const float* src, long long n, __half* hi, __half* lo, const float* absmax, float* inv_scale, cudaStream_t s) {
  split_flat_kernel<<<(unsigned)((n + TPB - 1) / TPB), TPB, 0, s>>>(src, n, hi, lo, absmax, inv_scale);
  SSNB_LAUNCH_CHECK("split_flat_kernel");
  return 0;
}
int launch_split_all(const SplitTable& t, int total_blocks, cudaStream_t s) {
  if (t.n <= 0) return 0;
  split_all_kernel<<<(unsigned)total_blocks, TPB, 0, s>>>(t);
  SSNB_LAUNCH_CHECK("split_all_kernel");
  return 0;
}
int launch_planes_to_nchw(View planes, int F, float scale, float* dst, cudaStream_t s) {
  const long long n = (long long)F * planes.C * planes.H * planes.W;
  planes_to_nchw_kernel<<<(unsigned)((n + TPB - 1) / TPB), TPB, 0, s>>>((const __half*)planes.base, planes.lo_off, F, planes.C, planes.H, planes.W,
                                                                      planes.pitch, planes.coff, scale, dst);
  SSNB_LAUNCH_CHECK("planes_to_nchw_kernel");
  return 0;
}
int launch_nchw_to_s2d_split(const float* src, int F, int Cin, int H, int W, __half* dst_hi, long long lo_off, int Cs, cudaStream_t s) {
  const long long n = (long long)F * (H / 2) * (W / 2) * 4;
  const unsigned g = (unsigned)((n + TPB - 1) / TPB);
  if (Cs == 16 && Cin == 3) nchw_to_s2d_split_kernel<16, 3><<<g, TPB, 0, s>>>(src, F, Cin, H, W, dst_hi, lo_off);
  else if (Cs == 40 && Cin == 10) nchw_to_s2d_split_kernel<40, 10><<<g, TPB, 0, s>>>(src, F, Cin, H, W, dst_hi, lo_off);
  else if (Cs == 16) nchw_to_s2d_split_kernel<16, 0><<<g, TPB, 0, s>>>(src, F, Cin, H, W, dst_hi, lo_off);
  else if (Cs == 40) nchw_to_s2d_split_kernel<40, 0><<<g, TPB, 0, s>>>(src, F, Cin, H, W, dst_hi, lo_off);
  else { set_thread_error("nchw_to_s2d_split: unsupported channel count (RGB 3 or Flow 10)"); return 1; }
  SSNB_LAUNCH_CHECK("nchw_to_s2d_split_kernel");
  return 0;
}

}  // namespace ssnb
