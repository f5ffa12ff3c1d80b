// fmha_sm100.cu -- tcgen05 / TMEM / TMA flash-attention forward (prefill + DiT joint attention).
// Placeholder translation unit: the kernel lands in a later commit; until then every call is
// routed to sdpa_generic by the dispatcher.
#include "omx_common.cuh"
#include "omx_internal.h"

namespace omx {

bool fmha_sm100_supported(const SdpaArgs&, const char** why) {
  if (why) *why = "tcgen05 kernel not built yet";
  return false;
}

void fmha_sm100(const SdpaArgs&, cudaStream_t) { OMX_CHECK(false, "fmha_sm100: not available"); }

}  // namespace omx
