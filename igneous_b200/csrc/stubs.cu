// stubs.cu -- entry points declared in include/igneous_b200.h whose kernels are
// not written yet.  They fail loudly (IGN_ERR_UNSUPPORTED); there is no CPU
// fallback anywhere in this library.  Entries move out of this file as the
// corresponding .cu lands.
#include "common.cuh"
using namespace ign;
#define IGN_STUB(name) \
  { set_error(name ": not implemented yet"); return IGN_ERR_UNSUPPORTED; }
extern "C" {
int ign_group_unique_id(void*) IGN_STUB("ign_group_unique_id")
int ign_group_init(ign_ctx*, int, int, const void*, ign_group**) IGN_STUB("ign_group_init")
int ign_group_destroy(ign_group*) IGN_STUB("ign_group_destroy")
int ign_ccl6_sharded_dev(ign_group*, const void*, int, uint64_t, uint64_t, uint64_t, const void*, uint64_t*, uint64_t*) IGN_STUB("ign_ccl6_sharded_dev")
}
