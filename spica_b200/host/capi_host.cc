// C entry point over the host for bindings that cannot hold C++ objects (ctypes in tests/ and
// bench.py): load a scene file, render it on N GPUs, return the normalised image.
#include <cstring>

#include "core.h"

using namespace spica;

extern "C" {

// out_rgb: height*width*3 floats (may be NULL to only query the size). Returns 0, or 1 when the
// buffer is too small (width / height are still written). Errors abort, like the reference.
int sph_render_scene(const char* xml_path, const char* output_prefix, int gpus, unsigned long long seed, int spp_override,
                     float* out_rgb, long long out_capacity, int* width, int* height) {
    RenderParams& params = RenderParams::getInstance();
    params.clear();
    params.add("numUserThreads", 1);
    params.add("outputFile", std::string(output_prefix ? output_prefix : "/tmp/spica_host_out"));
    HostOptions& opt = hostOptions();
    opt.gpus = gpus > 0 ? gpus : 1;
    opt.seed = seed;
    opt.savePasses = false;
    opt.sppOverride = spp_override;
    Image result;
    opt.onImage = [&](const Image& img) { result = img; };
    SceneParser parser(xml_path);
    parser.parse();
    opt.onImage = nullptr;
    if (width) *width = result.width;
    if (height) *height = result.height;
    const long long need = (long long)result.width * result.height * 3;
    if (!out_rgb) return 0;
    if (out_capacity < need) return 1;
    for (long long i = 0; i < need; i++) out_rgb[i] = (float)result.rgb[(size_t)i];
    return 0;
}

// Host-side logic only (no device): parse the scene and report what would be sent over the C ABI.
// info[0..7] = {n_triangles, n_lights, width, height, sampleCount, maxDepth, n_emitter_triangles, filter kind};
// camera = camera_to_world[16] then raster_to_camera[16]; verts (may be NULL) = n_triangles*9 doubles.
int sph_parse_scene(const char* xml_path, long long info[8], double camera[32], double* verts, long long verts_capacity) {
    RenderParams& params = RenderParams::getInstance();
    params.clear();
    params.add("numUserThreads", 1);
    params.add("outputFile", std::string("/tmp/spica_host_parse"));
    SceneParser parser(xml_path);
    parser.load();
    const auto& prims = parser.primitives();
    const auto cam = parser.camera();
    long long emit = 0;
    for (const auto& p : prims) emit += p->light ? 1 : 0;
    info[0] = (long long)prims.size(); info[1] = (long long)parser.lights().size();
    info[2] = cam->film->width(); info[3] = cam->film->height();
    info[4] = params.getInt("sampleCount", -1); info[5] = params.getInt("maxDepth", -1);
    info[6] = emit; info[7] = cam->film->filter()->kind();
    std::memcpy(camera, cam->cameraToWorld.getMat().m, sizeof(double) * 16);
    std::memcpy(camera + 16, cam->rasterToCamera.getMat().m, sizeof(double) * 16);
    if (verts) {
        if (verts_capacity < (long long)prims.size() * 9) return 1;
        for (size_t i = 0; i < prims.size(); i++) for (int k = 0; k < 3; k++) {
            verts[i * 9 + k * 3] = prims[i]->tri.p[k].x; verts[i * 9 + k * 3 + 1] = prims[i]->tri.p[k].y; verts[i * 9 + k * 3 + 2] = prims[i]->tri.p[k].z;
        }
    }
    return 0;
}

}  // extern "C"
