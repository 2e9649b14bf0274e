// bfb_fit.cu -- placeholder (replaced by the Gram / Cholesky implementation)
#include "bfb_common.cuh"
void bfb_fit_free(bfb_context *) {}
extern "C" int bfb_fit_begin(bfb_handle) { bfb_set_error("fit not built yet"); return BFB_ERR_STATE; }
