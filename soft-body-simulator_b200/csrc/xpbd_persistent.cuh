// Persistent, region-resident schedule (one cooperative kernel per substep).
#pragma once

#include "scene_build.h"
#include "xpbd_kernels.cuh"

#include <stdexcept>

namespace sbsb200 {

template <typename R>
struct PersistentPlan
{
    static int32_t regions_for(int sm_count, int64_t /*n_tets*/) { return sm_count; }
    void build(HostScene const&, ColourClass const&, RegionPlan const&, DeviceScene<R> const&, cudaStream_t, int)
    {
        throw std::runtime_error("persistent schedule not built into this library yet");
    }
    int64_t substep(DeviceScene<R> const&, R, int, bool, bool, cudaStream_t) { return 0; }
};

} // namespace sbsb200
