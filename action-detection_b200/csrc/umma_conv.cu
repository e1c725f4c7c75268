// placeholder: tcgen05 path not wired yet (all layers run on the SIMT engine)
#include "umma_conv.cuh"
namespace ssnb {
void umma_context_init(UmmaContext& ctx, bool fp16) { ctx.active = false; (void)fp16; }
void umma_context_destroy(UmmaContext&) {}
void umma_plan_workspace(UmmaContext&, size_t&) {}
int umma_conv_bind(UmmaContext&, UmmaConvPlan& p, View, View, int, int, int, int, int, int, char*, int, const float*) { p.enabled = false; return 0; }
int umma_conv_pack(UmmaContext&, UmmaConvPlan&, const __half*, int, int, int, cudaStream_t) { return 0; }
int umma_conv_forward(UmmaContext&, const UmmaConvPlan&, cudaStream_t) { return 1; }
}  // namespace ssnb
