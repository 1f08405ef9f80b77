// sp_slab.cu — slab decomposition over the GPUs of one node (placeholder until the NCCL path lands).
#include "sp_internal.cuh"

struct SlabState {
    int rank = 0, nranks = 1;
};

void sp_slab_free(sp_system* s) {
    delete s->slab;
    s->slab = nullptr;
}
int sp_slab_allreduce_device(sp_system* s, double*, int, int) {
    if (!s->slab) return SP_OK;
    return sp_fail(s, SP_ERR_STATE, "slab all-reduce not available in this build");
}

extern "C" {
int32_t sp_slab_unique_id(uint8_t id[128]) {
    (void)id;
    return sp_fail(nullptr, SP_ERR_STATE, "slab decomposition not available in this build");
}
int32_t sp_slab_init(sp_system* s, const uint8_t*, int32_t, int32_t, int32_t, int32_t) {
    return sp_fail(s, SP_ERR_STATE, "slab decomposition not available in this build");
}
int32_t sp_slab_create_cell_list(sp_system* s) { return sp_fail(s, SP_ERR_STATE, "not a slab system"); }
int32_t sp_slab_halo_refresh(sp_system* s, const int32_t*, int32_t) { return sp_fail(s, SP_ERR_STATE, "not a slab system"); }
int32_t sp_slab_num_owned(sp_system* s, int64_t*) { return sp_fail(s, SP_ERR_STATE, "not a slab system"); }
int32_t sp_slab_allreduce(sp_system* s, double*, int32_t, int32_t) { return sp_fail(s, SP_ERR_STATE, "not a slab system"); }
}
