// Concrete plugin classes of the host (registered under the reference's plugin type names).
#pragma once
#include "core.h"

namespace spica {

class AreaLight : public Light {            // lights/area.cc
public:
    explicit AreaLight(const Spectrum& Lemit) : Lemit(Lemit) {}
    int kind() const override { return SPB_LIGHT_AREA; }
    Spectrum Lemit;
};

class Envmap : public Light {               // lights/envmap.cc:49-55
public:
    int kind() const override { return SPB_LIGHT_ENVMAP; }
    Image image;
    Transform toWorld;
    double scale = 1.0, worldRadius = 2.0;
    Point3d worldCenter;
};

class CounterSampler : public Sampler {     // samplers/{independent,ldsampler,halton}.cc -> the counter-based stream
public:
    explicit CounterSampler(uint64_t seed = 0) : seed_(seed) {}
    double get1D() override;
    std::unique_ptr<Sampler> clone(unsigned int seed) const override { return std::make_unique<CounterSampler>(seed); }
private:
    uint64_t seed_, counter_ = 0;
};

}  // namespace spica
