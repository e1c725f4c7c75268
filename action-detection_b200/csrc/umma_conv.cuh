// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for SSNB_FAST_FP16 — interface.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cuda.h>

#include "common.cuh"

namespace ssnb {

struct UmmaContext {
  bool active = false;
  void* encode_tiled = nullptr;     // cuTensorMapEncodeTiled, resolved through cudaGetDriverEntryPoint
  size_t ws_base = 0, ws_bytes = 0; // region of the engine workspace owned by the tcgen05 path
  int num_sms = 148;
};

struct UmmaConvPlan {
  bool enabled = false;
  CUtensorMap tmap_a, tmap_b;
  int F = 0, H = 0, W = 0, Cin = 0, Cout = 0, k = 1, pad = 0;
  int bw = 0, bh = 0, bf = 0;          // TMA box (pixels) = bw*bh*bf = 128 rows of the M tile
  int tiles_w = 0, tiles_h = 0, tiles_f = 0, n_tiles = 1, block_n = 0, kchunks = 0;
  void* out = nullptr; int out_pitch = 0, out_coff = 0;
  const float* bias = nullptr;
  size_t wpack_off = 0;                // packed fp16 weights [tap][Cout_pad][Cin_pad]
  int cin_pad = 0, cout_pad = 0;
};

void umma_context_init(UmmaContext& ctx, bool fp16);
void umma_context_destroy(UmmaContext& ctx);
void umma_plan_workspace(UmmaContext& ctx, size_t& off);
int umma_conv_bind(UmmaContext& ctx, UmmaConvPlan& p, View in, View out, int F, int cin, int cout, int k, int stride,
                   int pad, char* ws, int conv_idx, const float* bias);
int umma_conv_pack(UmmaContext& ctx, UmmaConvPlan& p, const __half* wf, int cin, int cout, int k, cudaStream_t s);
int umma_conv_forward(UmmaContext& ctx, const UmmaConvPlan& p, cudaStream_t s);

}  // namespace ssnb
