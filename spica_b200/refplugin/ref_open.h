// Opens the reference's class declarations for the plugin shim.
//
// plugins/path.so and plugins/bvh.so built from this directory are loaded by the UNMODIFIED
// reference host (`spica -i scene.xml`) through its own loader (core/cobject.cc:32-50:
// dlopen("plugins/<type>.so") + dlsym("createInstance")).  They are compiled against the reference's
// headers where they lie ($(REF)/sources, never copied) and run inside the reference process, so
// they share its libspica_core.so.  The scene the reference parsed is held in objects whose
// parameters are private fields (bsdfs/diffuse.h:25-27, lights/area.h:52-53, core/camera.h:54-60 ...);
// reading them is the only way to resolve the scene into the PODs of include/spica_b200.h, so the
// access specifiers are opened here -- around the reference headers only, after every standard
// header they use has already been included (libstdc++'s <sstream> does not survive the define).
// Access specifiers do not change layout or mangling under the Itanium ABI.
#ifndef SPICA_B200_REF_OPEN_H_
#define SPICA_B200_REF_OPEN_H_

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <stack>
#include <string>
#include <thread>
#include <tuple>
#include <type_traits>
#include <typeinfo>
#include <unordered_map>
#include <utility>
#include <vector>

#define private public
#define protected public
#include "core/common.h"
#include "core/cobject.h"
#include "core/renderparams.h"
#include "core/accelerator.h"
#include "core/bounds3d.h"
#include "core/camera.h"
#include "core/constant.h"
#include "core/film.h"
#include "core/filter.h"
#include "core/image.h"
#include "core/integrator.h"
#include "core/interaction.h"
#include "core/light.h"
#include "core/material.h"
#include "core/mipmap.h"
#include "core/parallel.h"
#include "core/primitive.h"
#include "core/ray.h"
#include "core/sampler.h"
#include "core/scene.h"
#include "core/shape.h"
#include "core/spectrum.h"
#include "core/texture.h"
#include "core/transform.h"
#include "core/triangle.h"
// plugin headers carry their factory in the header (bsdfs/diffuse.h:30, accelerators/bvh.h:97): only the
// class layouts are wanted here, the shim exports its own createInstance
#undef SPICA_EXPORT_PLUGIN
#define SPICA_EXPORT_PLUGIN(name, descr)
#undef SPICA_EXPORT_ACCEL_PLUGIN
#define SPICA_EXPORT_ACCEL_PLUGIN(name, descr)
#include "bsdfs/conductor.h"
#include "bsdfs/dielectric.h"
#include "bsdfs/diffuse.h"
#include "bsdfs/plastic.h"
#include "bsdfs/roughconductor.h"
#include "bsdfs/roughdielectric.h"
#include "bsdfs/roughplastic.h"
#include "cameras/perspective.h"
#include "filters/box.h"
#include "filters/gaussian.h"
#include "filters/tent.h"
#include "lights/area.h"
#include "lights/envmap.h"
#include "textures/bitmap.h"
#include "textures/checkerboard.h"
#undef private
#undef protected

#endif  // SPICA_B200_REF_OPEN_H_
