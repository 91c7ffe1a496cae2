// Mitsuba-0.5-style scene loading with the reference's element semantics
// (spica/sceneparser.cc:58-334): depth-first walk, children stored before their parent is built,
// parameters kept in one RenderParams bag, objects made through the plugin manager by type name.
#include <new>
#include <cstring>
#include <ctime>
#include <filesystem>

#include "core.h"
#include "plugins.h"
#include "xml.h"

namespace fs = std::filesystem;

namespace spica {

namespace {
std::vector<std::string> split(const std::string& str, const std::string& delim) {   // sceneparser.cc:25-35
    size_t prev = 0, cur;
    std::vector<std::string> ret;
    while ((cur = str.find_first_of(delim, prev)) != std::string::npos) { ret.push_back(str.substr(prev, cur - prev)); prev = cur + delim.size(); }
    ret.push_back(str.substr(prev));
    return ret;
}
double str2double(const std::string& s) {
    char* ep; const double r = strtod(s.c_str(), &ep);
    if (strlen(ep) != 0) Warning("Following part could not be parsed as double: %s\n", ep);
    return r;
}
const char* getAttribute(const xml::Element* e, const char* name) {
    const char* v = e->attribute(name);
    if (!v) FatalError("Element \"%s\" does not have attribute \"%s\"", e->name.c_str(), name);
    return v;
}
struct ShapeMarker : CObject {};
}  // namespace

struct SceneParser::Impl {
    std::string xmlFile;
    RenderParams& params = RenderParams::getInstance();
    PluginManager& plugins = PluginManager::getInstance();
    std::shared_ptr<Camera> camera;
    std::vector<std::shared_ptr<Primitive>> primitives;
    std::vector<std::shared_ptr<Light>> lights;
    bool waitAreaLight = false;

    Transform parseTransform(const xml::Element* parent) {              // sceneparser.cc:120-166
        Transform trans;
        for (const auto& ch : parent->children) {
            const xml::Element* e = ch.get();
            Transform sub;
            if (e->name == "matrix") {
                const auto vals = split(getAttribute(e, "value"), " ");
                SpicaAssert(vals.size() == 16, "# of matrix values is not 16!");
                double m[4][4];
                for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[i][j] = str2double(vals[i * 4 + j]);
                sub = Transform(Matrix4x4(m));
            } else if (e->name == "scale") {
                sub = Transform::scale(str2double(getAttribute(e, "x")), str2double(getAttribute(e, "y")), str2double(getAttribute(e, "z")));
            } else if (e->name == "rotate") {
                const Vector3d ax(str2double(getAttribute(e, "x")), str2double(getAttribute(e, "y")), str2double(getAttribute(e, "z")));
                sub = Transform::rotate(str2double(getAttribute(e, "angle")), ax);      // radians, like the reference
            } else if (e->name == "translate") {
                sub = Transform::translate(Vector3d(str2double(getAttribute(e, "x")), str2double(getAttribute(e, "y")), str2double(getAttribute(e, "z"))));
            } else if (e->name == "lookAt") {
                sub = Transform::lookAt(Vector3d(getAttribute(e, "origin")), Vector3d(getAttribute(e, "target")), Vector3d(getAttribute(e, "up")));
            }
            trans = sub * trans;
        }
        return trans;
    }

    std::shared_ptr<Primitive> createPrimitive(const Triangle& tri, const Transform& transform, const std::shared_ptr<SurfaceMaterial>& material) {
        std::shared_ptr<Light> light;
        if (waitAreaLight) {                                                // sceneparser.cc:172-179: one AreaLight per triangle
            params.add("shape", std::static_pointer_cast<CObject>(std::make_shared<ShapeMarker>()));
            params.add("toWorld", transform);
            plugins.initModule("area");
            light = std::shared_ptr<Light>((Light*)plugins.createObject("area", params));
            lights.push_back(light);
        }
        auto p = std::make_shared<Primitive>();
        p->tri = tri; p->material = material; p->light = light;
        return p;
    }

    void parseChildren(const xml::Element* parent) {                        // sceneparser.cc:108-118
        for (const auto& ch : parent->children) {
            const xml::Element* e = ch.get();
            if (!e->noChildren() && e->name != "transform") parseChildren(e);
            storeToParam(e);
        }
    }

    void storeToParam(const xml::Element* e) {                              // sceneparser.cc:190-334
        const std::string node = e->name;
        std::string name = e->attribute("name") ? e->attribute("name") : "";
        if (node == "boolean") {
            // the reference tests `name == "boolean"` and then crashes on <boolean> elements
            // (sceneparser.cc:199,303); this host accepts them.
            const bool v = strcmp(getAttribute(e, "value"), "true") == 0;
            if (!name.empty()) params.add(name, v);
        } else if (node == "integer") {
            const int v = atoi(getAttribute(e, "value"));
            if (!name.empty()) params.add(name, v);
        } else if (node == "float") {
            const double v = strtod(getAttribute(e, "value"), nullptr);
            if (!name.empty()) params.add(name, v);
        } else if (node == "string") {
            std::string v = getAttribute(e, "value");
            if (name == "filename") {
                const fs::path base = fs::path(xmlFile).parent_path();
                std::error_code ec;
                const fs::path full = fs::canonical(base / fs::path(v), ec);
                if (ec) FatalError("Failed to open file: %s", (base / fs::path(v)).string().c_str());
                v = full.string();
            }
            if (!name.empty()) params.add(name, v);
        } else if (node == "rgb") {
            const Vector3d v(getAttribute(e, "value"));
            if (!name.empty()) params.add(name, Spectrum(v.x, v.y, v.z));
        } else if (node == "point" || node == "vector") {
            if (!name.empty()) {
                if (e->attribute("value")) params.add(name, Vector3d(std::string(e->attribute("value"))));
                else params.add(name, Vector3d(str2double(getAttribute(e, "x")), str2double(getAttribute(e, "y")), str2double(getAttribute(e, "z"))));
            }
        } else if (node == "spectrum") {
            const std::string s = getAttribute(e, "value");
            if (s.find(':') != std::string::npos) FatalError("wavelength:value spectra are outside this host's scope (use <rgb>): %s", s.c_str());
            const Vector3d v(s);
            if (!name.empty()) params.add(name, Spectrum(v.x, v.y, v.z));
        } else if (node == "transform") {
            const Transform t = parseTransform(e);
            if (!name.empty()) params.add(name, t);
        } else if (node == "shape") {
            const std::string type = getAttribute(e, "type");
            auto surface = std::static_pointer_cast<SurfaceMaterial>(params.getObject("bsdf", nullptr, true));
            if (params.getObject("subsurface", nullptr, true)) FatalError("subsurface materials are outside this host's scope");
            if (params.getObject("medium", nullptr, true)) FatalError("participating media are outside this host's scope");
            const Transform transform = params.getTransform("toWorld", Transform(), true);
            std::vector<Triangle> objTris; meshio::TriangleBuffer plyTris;
            const Triangle* tris = nullptr; size_t nTris = 0;
            if (type == "obj") { objTris = meshio::loadOBJ(params.getString("filename"), transform); tris = objTris.data(); nTris = objTris.size(); }
            else if (type == "ply") { plyTris = meshio::loadPLY(params.getString("filename"), transform); tris = plyTris.p; nTris = plyTris.n; }
            else FatalError("Failed to load plugin: plugins/%s.so (analytic shapes are outside this host's scope; use obj / ply)", type.c_str());
            if (nTris >= 4096) hostPhaseLap("mesh file read");
            if (waitAreaLight || nTris < 4096) {
                for (size_t i = 0; i < nTris; i++) primitives.push_back(createPrimitive(tris[i], transform, surface));
            } else {
                // a large mesh: its primitives in ONE allocation, constructed by all cores and handed out through aliasing pointers
                // (the plugin surface wants a std::shared_ptr<Primitive> per triangle, core/cobject.h:66-75; ten million
                // make_shared calls, or one thread touching 2.3 GB first, cost seconds)
                Primitive* raw = (Primitive*)std::malloc(sizeof(Primitive) * nTris);
                if (!raw) FatalError("out of memory for %zu primitives", nTris);
                parallelFor(nTris, [&](size_t b, size_t e) { for (size_t i = b; i < e; i++) { Primitive* q = new (raw + i) Primitive(); q->tri = tris[i]; q->material = surface; } });
                const size_t count = nTris;
                std::shared_ptr<Primitive> block(raw, [count](Primitive* q) {       // (may run during static destruction: no threads here)
                    for (size_t i = 0; i < count; i++) q[i].~Primitive();
                    std::free(q);
                });
                primitives.reserve(primitives.size() + nTris);
                for (size_t i = 0; i < nTris; i++) primitives.emplace_back(block, raw + i);
            }
            if (nTris >= 4096) hostPhaseLap("primitives created");
            waitAreaLight = false;
        } else if (node == "ref") {
            const std::string id = getAttribute(e, "id");
            auto obj = params.getObject(id);
            if (dynamic_cast<SurfaceMaterial*>(obj.get())) params.add("bsdf", obj);
        } else if (node == "integrator") {
            const std::string type = getAttribute(e, "type");
            SpicaAssert(!type.empty(), "Integrator type is not specified!");
            params.add("integrator", type);
        } else {
            const std::string type = getAttribute(e, "type");
            if (node == "emitter" && type == "area") { waitAreaLight = true; return; }
            SpicaAssert(!type.empty(), "Type parameter is not specified for \"%s\"", node.c_str());
            plugins.initModule(type);
            auto value = std::shared_ptr<CObject>(plugins.createObject(type, params));
            if (!name.empty()) params.add(name, value);
            else {
                const std::string id = e->attribute("id") ? e->attribute("id") : "";
                if (!id.empty()) params.add(id, value);
                params.add(node, value);
            }
            if (node == "sensor") {
                SpicaAssert(!camera, "Multiple cameras are specified!");
                camera = std::static_pointer_cast<Camera>(value);
            } else if (node == "emitter") {
                lights.push_back(std::static_pointer_cast<Light>(value));
            }
        }
    }
};

SceneParser::SceneParser(const std::string& xmlFile) : impl_(std::make_shared<Impl>()) {
    registerBuiltinPlugins();
    // defaults injected by the reference's parser (sceneparser.cc:58-65)
    impl_->params.add("maxDepth", 16);
    impl_->params.add("accelerator", std::string("bvh"));
    std::error_code ec;
    const fs::path p = fs::canonical(fs::absolute(fs::path(xmlFile)), ec);
    if (ec) FatalError("Failed to open file:%s\n", xmlFile.c_str());
    impl_->xmlFile = p.string();
}

void SceneParser::parse() { load(); render(); }
const std::vector<std::shared_ptr<Primitive>>& SceneParser::primitives() const { return impl_->primitives; }
const std::vector<std::shared_ptr<Light>>& SceneParser::lights() const { return impl_->lights; }
std::shared_ptr<Camera> SceneParser::camera() const { return impl_->camera; }

void SceneParser::load() {                                                  // sceneparser.cc:72-85
    Impl& I = *impl_;
    std::string err;
    auto root = xml::parseFile(I.xmlFile, &err);
    if (!root) FatalError("Failed to open file:%s (%s)\n", I.xmlFile.c_str(), err.c_str());
    SpicaAssert(root->name == "scene", "XML root node should be \"scene\"!");
    printf("Version: %s\n", root->attribute("version") ? root->attribute("version") : "(null)");
    I.parseChildren(root.get());
    SpicaAssert(I.camera != nullptr, "Sensor is not specified!");
}

void SceneParser::render() {                                                // sceneparser.cc:87-105
    Impl& I = *impl_;
    const std::string integType = I.params.getString("integrator");
    I.plugins.initModule(integType);
    auto integrator = std::shared_ptr<Integrator>((Integrator*)I.plugins.createObject(integType, I.params));
    const std::string accelType = I.params.getString("accelerator");
    auto accelerator = std::shared_ptr<Accelerator>(I.plugins.createAccelerator(accelType, I.primitives, I.params));
    Scene scene(accelerator, I.lights);
    MsgInfo("   Scene: %s", I.xmlFile.c_str());
    MsgInfo("    GPUs: %d", hostOptions().gpus);
    integrator->render(I.camera, scene, I.params);
}

}  // namespace spica
