// plugins/bvh.so for the UNMODIFIED reference host: replaces accelerators/bvh.cc behind the same
// factory symbols (core/cobject.h:66-75).  The tree lives in HBM (spb_bvh_build); see gpu_scene.h.
#include "gpu_scene.h"

extern "C" {
spica::CObject* createInstance(const std::vector<std::shared_ptr<spica::Primitive>>& primitives, spica::RenderParams& params) {
    return (spica::CObject*)(new spica::b200::GpuBVHAccel(primitives, params));
}
const char* getDescription() { return "B200 bounding volume hierarchy (8-wide compressed, resident in HBM)"; }
}
