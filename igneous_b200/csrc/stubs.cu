// stubs.cu -- entry points declared in include/igneous_b200.h whose kernels are
// not written yet.  They fail loudly (IGN_ERR_UNSUPPORTED); there is no CPU
// fallback anywhere in this library.  Entries move out of this file as the
// corresponding .cu lands.
#include "common.cuh"
using namespace ign;
#define IGN_STUB(name) \
  { set_error(name ": not implemented yet"); return IGN_ERR_UNSUPPORTED; }
extern "C" {
int ign_renumber(ign_ctx*, const void*, int, uint64_t, uint32_t*, uint64_t*, uint64_t, uint64_t*) IGN_STUB("ign_renumber")
int ign_renumber_dev(ign_ctx*, const void*, int, uint64_t, uint32_t*, uint64_t*, uint64_t, uint64_t*) IGN_STUB("ign_renumber_dev")
int ign_remap(ign_ctx*, void*, int, uint64_t, const uint64_t*, const uint64_t*, uint64_t, int) IGN_STUB("ign_remap")
int ign_remap_dev(ign_ctx*, void*, int, uint64_t, const uint64_t*, const uint64_t*, uint64_t, int) IGN_STUB("ign_remap_dev")
int ign_unique(ign_ctx*, const void*, int, uint64_t, uint64_t*, uint64_t*, uint64_t, uint64_t*) IGN_STUB("ign_unique")
int ign_mask(ign_ctx*, void*, int, uint64_t, const uint64_t*, uint64_t, int, uint64_t) IGN_STUB("ign_mask")
int ign_inverse_component_map(ign_ctx*, const void*, const void*, int, uint64_t, uint64_t*, uint64_t*) IGN_STUB("ign_inverse_component_map")
int ign_cast_dev(ign_ctx*, const void*, int, void*, int, uint64_t) IGN_STUB("ign_cast_dev")
int ign_mesh_begin(ign_ctx*, const void*, int, uint64_t, uint64_t, uint64_t, ign_mesher**) IGN_STUB("ign_mesh_begin")
int ign_mesh_begin_dev(ign_ctx*, const void*, int, uint64_t, uint64_t, uint64_t, ign_mesher**) IGN_STUB("ign_mesh_begin_dev")
int ign_mesh_num_ids(ign_mesher*, uint64_t*) IGN_STUB("ign_mesh_num_ids")
int ign_mesh_ids(ign_mesher*, uint64_t*, uint64_t) IGN_STUB("ign_mesh_ids")
int ign_mesh_counts(ign_mesher*, uint64_t, uint64_t*, uint64_t*) IGN_STUB("ign_mesh_counts")
int ign_mesh_totals(ign_mesher*, uint64_t*, uint64_t*) IGN_STUB("ign_mesh_totals")
int ign_mesh_get(ign_mesher*, uint64_t, const float*, int, float, int, float*, uint32_t*, uint64_t*, uint64_t*) IGN_STUB("ign_mesh_get")
int ign_mesh_free(ign_mesher*) IGN_STUB("ign_mesh_free")
int ign_group_unique_id(void*) IGN_STUB("ign_group_unique_id")
int ign_group_init(ign_ctx*, int, int, const void*, ign_group**) IGN_STUB("ign_group_init")
int ign_group_destroy(ign_group*) IGN_STUB("ign_group_destroy")
int ign_ccl6_sharded_dev(ign_group*, const void*, int, uint64_t, uint64_t, uint64_t, const void*, uint64_t*, uint64_t*) IGN_STUB("ign_ccl6_sharded_dev")
}
