// Stage profiling for the C-ABI (vlo_set_profiling / vlo_get_stage_times).
#include "vlo_internal.cuh"

static const char *kStageNames[ST_COUNT] = { "k0_organise", "k1_extract", "k1b_compact", "k2_grid_build", "k3_to_end", "k3_assoc",
                                             "k3_gn", "k5_register_coop", "k5_lin", "k5_assoc", "k6_imu", "k7_stack_ds",
                                             "k7_map_insert" };

void vlo_prof_begin(vlo_handle *h, int stage)
{
    if (!h->prof_enabled) return;
    if (h->prof_used + 2 > h->prof_events.size()) {
        size_t old = h->prof_events.size();
        h->prof_events.resize(old + 512);
        for (size_t k = old; k < h->prof_events.size(); k++) cudaEventCreate(&h->prof_events[k]);
    }
    h->prof_stage.push_back(stage);
    cudaEventRecord(h->prof_events[h->prof_used], h->stream);
}

void vlo_prof_end(vlo_handle *h, int stage)
{
    (void)stage;
    if (!h->prof_enabled) return;
    cudaEventRecord(h->prof_events[h->prof_used + 1], h->stream);
    h->prof_used += 2;
}

extern "C" int vlo_set_profiling(vlo_handle *h, int enable)
{
    if (!h) return VLO_ERR_INVALID_ARG;
    cudaStreamSynchronize(h->stream);
    h->prof_enabled = enable ? 1 : 0; h->prof_used = 0; h->prof_stage.clear();
    return VLO_OK;
}

extern "C" int vlo_stage_count(void) { return ST_COUNT; }
extern "C" const char *vlo_stage_name(int stage) { return (stage >= 0 && stage < ST_COUNT) ? kStageNames[stage] : ""; }

// ms[ST_COUNT], launches[ST_COUNT]: accumulated since the last call (or since profiling was enabled)
extern "C" int vlo_get_stage_times(vlo_handle *h, float *ms, int *launches)
{
    if (!h || !ms || !launches) return VLO_ERR_INVALID_ARG;
    VLO_CUDA(cudaStreamSynchronize(h->stream));
    for (int s = 0; s < ST_COUNT; s++) { ms[s] = 0.f; launches[s] = 0; }
    for (size_t k = 0; k * 2 < h->prof_used; k++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, h->prof_events[2 * k], h->prof_events[2 * k + 1]) == cudaSuccess) {
            ms[h->prof_stage[k]] += t; launches[h->prof_stage[k]] += 1;
        }
    }
    h->prof_used = 0; h->prof_stage.clear();
    return VLO_OK;
}

extern "C" void *vlo_stream(vlo_handle *h) { return h ? (void *)h->stream : nullptr; }
