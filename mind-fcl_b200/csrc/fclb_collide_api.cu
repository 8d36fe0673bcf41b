// fclb_collide_api.cu -- C ABI entry points for batched fcl::collide and the
// direct GJK+EPA path.  (Kernels: fclb_collide_impl.cuh.)
#include "fclb_internal.h"

extern "C" {
int fclb_collide_batch_dev(fclb_handle, const fclb_pair*, const void*, const void*, size_t, int, const fclb_request*,
                           uint32_t, void*, uint32_t*) {
  return FCLB_ERR_UNSUPPORTED;
}
int fclb_collide_batch_host(fclb_handle, const fclb_pair*, const void*, const void*, size_t, int, const fclb_request*,
                            uint32_t, void*, uint32_t*) {
  return FCLB_ERR_UNSUPPORTED;
}
int fclb_gjk_epa_batch_dev(fclb_handle, const fclb_pair*, const void*, const void*, size_t, int, const fclb_request*,
                           int32_t*, int32_t*, void*) {
  return FCLB_ERR_UNSUPPORTED;
}
int fclb_gjk_epa_batch_host(fclb_handle, const fclb_pair*, const void*, const void*, size_t, int, const fclb_request*,
                            int32_t*, int32_t*, void*) {
  return FCLB_ERR_UNSUPPORTED;
}
}
