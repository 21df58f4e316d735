// b200_path_native.cpp -- a NATIVE Mitsuba 3 integrator plugin that forwards to libb200pt.so (C ABI, include/b200pt.h).
//
// This is the compiled form of the shim sketched in INTEGRATION.md section 2: a regular plugin
// (`MI_EXPORT_PLUGIN`, include/mitsuba/core/object.h:343-347; loaded by PluginManager through dlopen + init_plugin,
// src/core/plugin.cpp:126-149) whose class derives from SamplingIntegrator<Float, Spectrum> and overrides render()
// (include/mitsuba/render/integrator.h:443-448). It is built by native/build_shim.sh against the headers of the
// reference checkout and the runtime of oracle/_ref -- test infrastructure like oracle/_ref itself; the product is the C ABI.
//
// Scope of the native extractor: what mi.cornell_box()-class scenes contain -- triangle meshes (packed records incl. tangent
// frames and the FaceUVFlipped bit, which C++ can read directly: Mesh::packed_face), SmoothDiffuse BSDFs with a uniform rgb
// reflectance (optionally inside TwoSidedBRDF), uniform area lights (rectangles sampled through their to_world), a
// perspective sensor, box / gaussian reconstruction filter, the independent sampler. Everything else throws and points to
// the Python plugin (mitsuba3_b200/mitsuba_plugin.py), whose extractor covers the whole hot-path scope.
//
//   B200PT_LIB        path of libb200pt.so (default: the library next to the Python package, resolved through
//                     B200PT_ROOT or the working directory)
//   B200PT_SHIM_DUMP  if set: write the scene description the shim built to this file and return a black image without
//                     touching the GPU (tests/test_native_shim.py compares it with the Python extractor's, on a CPU box)
#include <mitsuba/core/properties.h>
#include <mitsuba/core/transform.h>
#include <mitsuba/render/bsdf.h>
#include <mitsuba/render/emitter.h>
#include <mitsuba/render/film.h>
#include <mitsuba/render/integrator.h>
#include <mitsuba/render/mesh.h>
#include <mitsuba/render/sampler.h>
#include <mitsuba/render/scene.h>
#include <mitsuba/render/sensor.h>

#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <regex>
#include <string>
#include <typeinfo>
#include <vector>

#include "b200pt.h"

NAMESPACE_BEGIN(mitsuba)

namespace {

// ---- the C ABI, bound at run time (no link-time dependency of the plugin on the CUDA library) ------------------------------
struct B200Api {
    void *handle = nullptr;
    uint32_t (*abi_version)() = nullptr;
    const char *(*last_error)() = nullptr;
    b200pt_status (*scene_create)(const b200pt_scene_desc *, int, b200pt_scene **) = nullptr;
    void (*scene_destroy)(b200pt_scene *) = nullptr;
    b200pt_status (*render)(b200pt_scene *, const b200pt_render_params *, float *) = nullptr;

    void load() {
        if (handle) return;
        std::string path;
        if (const char *e = getenv("B200PT_LIB")) path = e;
        else {
            const char *root = getenv("B200PT_ROOT");
            path = std::string(root ? root : ".") + "/mitsuba3_b200/lib/libb200pt.so";
        }
        handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!handle) Throw("b200_path_native: cannot load \"%s\": %s", path, dlerror());
        auto sym = [&](const char *n) { void *p = dlsym(handle, n); if (!p) Throw("b200_path_native: %s is not exported by %s", n, path); return p; };
        abi_version = (uint32_t (*)()) sym("b200pt_abi_version");
        last_error = (const char *(*)()) sym("b200pt_last_error");
        scene_create = (b200pt_status (*)(const b200pt_scene_desc *, int, b200pt_scene **)) sym("b200pt_scene_create");
        scene_destroy = (void (*)(b200pt_scene *)) sym("b200pt_scene_destroy");
        render = (b200pt_status (*)(b200pt_scene *, const b200pt_render_params *, float *)) sym("b200pt_render");
        if (abi_version() != B200PT_ABI_VERSION) Throw("b200_path_native: ABI version mismatch (library %u, header %u)", abi_version(), (uint32_t) B200PT_ABI_VERSION);
    }
};

// ---- parameters of a plugin through Object::traverse (object.h:399-436), flattened to "a.b.c" -> (pointer, type) --------------
struct Collector : TraversalCallback {
    struct Entry { void *ptr; const std::type_info *type; };
    std::map<std::string, Entry> values;
    std::map<std::string, Object *> objects;
    std::string prefix;
    void put_value(std::string_view name, void *value, uint32_t, const std::type_info &type) override {
        values[prefix + std::string(name)] = { value, &type };
    }
    void put_object(std::string_view name, Object *value, uint32_t) override {
        if (!value) return;
        std::string key = prefix + std::string(name);
        objects[key] = value;
        std::string saved = prefix;
        prefix = key + ".";
        value->traverse(this);
        prefix = saved;
    }
};

}  // namespace

template <typename Float, typename Spectrum>
class B200PathNative final : public SamplingIntegrator<Float, Spectrum> {
public:
    MI_IMPORT_BASE(SamplingIntegrator, m_hide_emitters)
    MI_IMPORT_TYPES(Scene, Sensor, Sampler, Medium, Mesh, Shape, BSDF, Emitter, Film, ReconstructionFilter)

    B200PathNative(const Properties &props) : Base(props) {
        // MonteCarloIntegrator (integrator.cpp:539-550): same keys, same checks
        m_max_depth = props.get<int>("max_depth", -1);
        if (m_max_depth < 0 && m_max_depth != -1)
            Throw("\"max_depth\" must be set to -1 (infinite) or a value >= 0");
        m_rr_depth = props.get<int>("rr_depth", 5);
        if (m_rr_depth <= 0)
            Throw("\"rr_depth\" must be set to a value greater than zero!");
        m_device = props.get<int>("device", 0);
    }

    ~B200PathNative() {
        if (m_handle && m_api.scene_destroy) m_api.scene_destroy(m_handle);
    }

    std::pair<Spectrum, Mask> sample(const Scene *, Sampler *, const RayDifferential3f &, const Medium *, Float *, Mask) const override {
        Throw("b200_path_native renders whole frames on the GPU: sample() is not available");
    }

    TensorXf render(Scene *scene, Sensor *sensor, UInt32 seed, uint32_t spp, bool develop, bool /* evaluate */) override {
        if (!develop)
            Throw("b200_path_native: develop=false is not supported");
        Description d;
        describe(scene, sensor, d);
        ScalarVector2u crop = sensor->film()->crop_size();
        size_t n = (size_t) crop.x() * crop.y() * 3;
        std::vector<float> img(n, 0.f);
        if (const char *dump = getenv("B200PT_SHIM_DUMP")) {
            write_dump(d, dump);
        } else {
            m_api.load();
            if (!m_handle || m_scene != scene) {
                if (m_handle) m_api.scene_destroy(m_handle);
                m_handle = nullptr;
                check(m_api.scene_create(&d.desc, m_device, &m_handle));
                m_scene = scene;
            }
            b200pt_render_params p;
            std::memset(&p, 0, sizeof(p));
            p.seed = (uint32_t) dr::slice(seed, 0);
            p.spp = spp ? spp : (uint32_t) sensor->sampler()->sample_count();
            p.max_depth = m_max_depth; p.rr_depth = m_rr_depth; p.hide_emitters = m_hide_emitters ? 1 : 0;
            p.shard_rank = 0; p.shard_count = 1; p.tile_size = 32;
            check(m_api.render(m_handle, &p, img.data()));
        }
        size_t shape[3] = { crop.y(), crop.x(), 3 };
        return TensorXf(dr::load<typename TensorXf::Array>(img.data(), n), 3, shape);
    }

    std::string to_string() const override {
        return tfm::format("B200PathNative[\n  max_depth = %d,\n  rr_depth = %d\n]", m_max_depth, m_rr_depth);
    }

    MI_DECLARE_CLASS(B200PathNative)

private:
    // Owns every array the POD description points to
    struct Description {
        b200pt_scene_desc desc;
        std::vector<b200pt_texture> textures;
        std::vector<b200pt_bsdf> bsdfs;
        std::vector<b200pt_shape> shapes;
        std::vector<b200pt_emitter> emitters;
        std::vector<std::vector<float>> vertex_data;
        std::vector<std::vector<uint32_t>> face_data;
        std::vector<std::string> names;
    };

    void check(b200pt_status st) const {
        if (st != B200PT_OK)
            Throw("b200pt: %s", m_api.last_error ? m_api.last_error() : "error");
    }

    static float scalar(const Float &v) { return (float) dr::slice(v, 0); }

    /// Uniform rgb value of a texture child ("<key>.value" of the traversed plugin); throws for anything else
    static void uniform_rgb(const Collector &c, const std::string &key, float out[3]) {
        auto it = c.values.find(key + ".value");
        if (it == c.values.end())
            Throw("b200_path_native: \"%s\" is not a uniform value (bitmap / spectrum textures: use the Python plugin)", key);
        const std::type_info &t = *it->second.type;
        if (t == typeid(Color<Float, 3>)) {
            const Color<Float, 3> &v = *(const Color<Float, 3> *) it->second.ptr;
            out[0] = scalar(v.x()); out[1] = scalar(v.y()); out[2] = scalar(v.z());
        } else if (t == typeid(Float)) {
            out[0] = out[1] = out[2] = scalar(*(const Float *) it->second.ptr);
        } else {
            Throw("b200_path_native: unsupported parameter type of \"%s.value\"", key);
        }
    }

    int add_const_texture(Description &d, const float rgb[3]) {
        b200pt_texture t;
        std::memset(&t, 0, sizeof(t));
        t.kind = B200PT_TEX_CONST; t.channels = 3; t.differentiable = 1;
        t.value[0] = rgb[0]; t.value[1] = rgb[1]; t.value[2] = rgb[2];
        t.to_uv[0] = t.to_uv[4] = t.to_uv[8] = 1.f;
        d.textures.push_back(t);
        return (int) d.textures.size() - 1;
    }

    int add_bsdf(Description &d, const BSDF *bsdf, std::map<const BSDF *, int> &seen) {
        auto it = seen.find(bsdf);
        if (it != seen.end()) return it->second;
        Collector c;
        const_cast<BSDF *>(bsdf)->traverse(&c);
        std::string cls(bsdf->class_name());
        std::string prefix;
        b200pt_bsdf b;
        std::memset(&b, 0, sizeof(b));
        for (int k = 0; k < B200PT_MAX_SLOTS; ++k) b.tex[k] = -1;
        if (cls == "TwoSidedBRDF") {
            auto ob = c.objects.find("brdf_0");
            if (ob == c.objects.end()) Throw("b200_path_native: twosided without a nested BRDF");
            cls = std::string(ob->second->class_name());
            prefix = "brdf_0.";
            b.twosided = 1;
        }
        if (cls != "SmoothDiffuse")
            Throw("b200_path_native: BSDF class %s: the native shim extracts diffuse materials only, use the Python plugin "
                  "(mitsuba3_b200.mitsuba_plugin.register) for the full hot-path scope", cls);
        float rgb[3];
        uniform_rgb(c, prefix + "reflectance", rgb);
        b.type = B200PT_BSDF_DIFFUSE;
        b.tex[B200PT_SLOT_REFLECTANCE] = add_const_texture(d, rgb);
        b.eta = 1.f;
        d.bsdfs.push_back(b);
        seen[bsdf] = (int) d.bsdfs.size() - 1;
        return seen[bsdf];
    }

    void describe(Scene *scene, Sensor *sensor, Description &d) {
        std::memset(&d.desc, 0, sizeof(d.desc));
        std::map<const BSDF *, int> seen;
        // emitter order = Scene::emitters() (scene.cpp:45-61): it decides which emitter a sample picks
        std::vector<const Emitter *> ems;
        for (auto &e : scene->emitters()) ems.push_back(e.get());
        d.emitters.resize(ems.size());
        for (auto &e : d.emitters) { std::memset(&e, 0, sizeof(e)); e.shape = -1; e.radiance_tex = -1; e.sampling_weight = 1.f; }
        std::vector<bool> em_done(ems.size(), false);

        for (auto &sref : scene->shapes()) {
            const Shape *shape = sref.get();
            const Mesh *mesh = dynamic_cast<const Mesh *>(shape);
            if (!mesh)
                Throw("b200_path_native: only triangle meshes are on the hot path (shape %s)", shape->id());
            uint32_t nv = (uint32_t) mesh->vertex_count(), nf = (uint32_t) mesh->face_count();
            b200pt_shape s;
            std::memset(&s, 0, sizeof(s));
            s.n_vertices = nv; s.n_faces = nf;
            // packed vertex records (mesh_utils.h:19-46) and packed face records incl. the flags lane
            d.vertex_data.emplace_back((size_t) nv * 8);
            {
                auto buf = mesh->packed_vertices();
                if constexpr (dr::is_jit_v<Float>) { dr::eval(buf); dr::sync_thread(); }
                std::memcpy(d.vertex_data.back().data(), buf.data(), (size_t) nv * 8 * sizeof(float));
            }
            d.face_data.emplace_back((size_t) nf * 4);
            for (uint32_t f = 0; f < nf; ++f) {
                auto rec = mesh->packed_face(ScalarUInt32(f));
                for (int k = 0; k < 4; ++k) d.face_data.back()[4 * (size_t) f + k] = (uint32_t) dr::slice(rec[k], 0);
                if (!mesh->packs_tangent()) d.face_data.back()[4 * (size_t) f + 3] = 0u;     // per-face BSDF indices: not on the hot path
                else d.face_data.back()[4 * (size_t) f + 3] &= 0x80000000u;                 // FaceUVFlipped (mesh_utils.h:32)
            }
            s.vertices = d.vertex_data.back().data(); s.faces = d.face_data.back().data();
            s.layout = (mesh->has_normals() ? B200PT_LAYOUT_NORMALS : 0) | (mesh->has_texcoords() ? B200PT_LAYOUT_TEXCOORDS : 0) |
                       (mesh->packs_tangent() ? B200PT_LAYOUT_TANGENTS : 0);
            s.bsdf = add_bsdf(d, shape->bsdf(), seen);
            s.emitter = -1; s.sampling = B200PT_SAMPLING_NONE;
            if (shape->is_emitter()) {
                const Emitter *em = shape->emitter();
                size_t slot = 0;
                while (slot < ems.size() && ems[slot] != em) ++slot;
                if (slot == ems.size()) Throw("b200_path_native: emitter of a shape is not in Scene::emitters()");
                Collector ec;
                const_cast<Emitter *>(em)->traverse(&ec);
                float rad[3];
                uniform_rgb(ec, "radiance", rad);
                b200pt_emitter &e = d.emitters[slot];
                e.shape = (int32_t) d.shapes.size(); e.radiance_tex = add_const_texture(d, rad); e.type = B200PT_EMITTER_AREA;
                auto sw = ec.values.find("sampling_weight");
                e.sampling_weight = sw != ec.values.end() && *sw->second.type == typeid(ScalarFloat) ? (float) *(const ScalarFloat *) sw->second.ptr : 1.f;
                em_done[slot] = true;
                s.emitter = (int32_t) slot;
                s.sampling = B200PT_SAMPLING_MESH;
                // Rectangle: sampled through its parameterisation (rectangle.cpp:159-179)
                Collector sc;
                const_cast<Shape *>(shape)->traverse(&sc);
                auto tw = sc.values.find("to_world");
                if (tw != sc.values.end() && nv == 4 && nf == 2) {
                    ScalarTransform4f t = to_scalar_transform(tw->second);
                    s.sampling = B200PT_SAMPLING_RECTANGLE;
                    for (int r = 0; r < 4; ++r)
                        for (int cidx = 0; cidx < 4; ++cidx)
                            s.to_world[4 * r + cidx] = (float) t.matrix(r, cidx);
                    ScalarNormal3f n = dr::normalize(t * ScalarNormal3f(0.f, 0.f, 1.f));
                    s.frame_n[0] = n.x(); s.frame_n[1] = n.y(); s.frame_n[2] = n.z();
                    ScalarVector3f du = t * ScalarVector3f(2.f, 0.f, 0.f), dv = t * ScalarVector3f(0.f, 2.f, 0.f);
                    s.inv_area = 1.f / dr::norm(dr::cross(du, dv));
                }
            }
            d.shapes.push_back(s);
        }
        for (size_t k = 0; k < ems.size(); ++k)
            if (!em_done[k])
                Throw("b200_path_native: emitter %s: only area lights are extracted natively, use the Python plugin", ems[k]->class_name());

        // ---- sensor / film -----------------------------------------------------------------------------------------------------
        Collector sc;
        sensor->traverse(&sc);
        auto need = [&](const char *k) -> const Collector::Entry & {
            auto it = sc.values.find(k);
            if (it == sc.values.end()) Throw("b200_path_native: the sensor exposes no \"%s\" (perspective sensors only)", k);
            return it->second;
        };
        float x_fov = scalar(*(const Float *) need("x_fov").ptr);
        const Film *film = sensor->film();
        ScalarVector2u size = film->size(), crop = film->crop_size();
        ScalarPoint2u off = film->crop_offset();
        float near_clip = (float) sensor_near(sensor), far_clip = (float) sensor_far(sensor);
        // in the variant's own arithmetic, like PerspectiveCamera::update_camera_transforms (perspective.cpp:150-170) and like
        // mi.perspective_projection in the Python extractor (JIT variants evaluate it with Dr.Jit's tan, not libm's)
        auto proj = perspective_projection<Float>(ScalarVector2i(size), ScalarVector2i(crop), ScalarVector2i(off), Float(x_fov), Float(near_clip), Float(far_clip));
        auto inv = proj.inverse().matrix;
        b200pt_sensor &se = d.desc.sensor;
        for (int r = 0; r < 4; ++r)
            for (int cidx = 0; cidx < 4; ++cidx)
                se.sample_to_camera[4 * r + cidx] = (float) dr::slice(inv(r, cidx), 0);
        {
            const auto &e = need("to_world");
            ScalarTransform4f tw = to_scalar_transform(e);
            for (int r = 0; r < 4; ++r)
                for (int cidx = 0; cidx < 4; ++cidx)
                    se.to_world[4 * r + cidx] = (float) tw.matrix(r, cidx);
        }
        se.near_clip = near_clip; se.far_clip = far_clip;
        se.film_size[0] = size.x(); se.film_size[1] = size.y(); se.crop_size[0] = crop.x(); se.crop_size[1] = crop.y();
        se.crop_offset[0] = off.x(); se.crop_offset[1] = off.y();
        const ReconstructionFilter *rf = film->rfilter();
        if (rf->is_box_filter()) { se.rfilter = B200PT_RFILTER_BOX; se.rfilter_stddev = 0.f; }
        else if (std::string(rf->class_name()) == "GaussianFilter") { se.rfilter = B200PT_RFILTER_GAUSSIAN; se.rfilter_stddev = (float) rf->radius() / 4.f; }
        else Throw("b200_path_native: rfilter %s is outside the hot-path scope", rf->class_name());
        const Sampler *sampler = sensor->sampler();
        if (std::string(sampler->class_name()) != "IndependentSampler")
            Throw("b200_path_native: sampler %s: only `independent` is on the hot path", sampler->class_name());
        {
            std::smatch m;
            std::string str = sampler->to_string();
            se.base_seed = std::regex_search(str, m, std::regex("base_seed\\s*=\\s*(\\d+)")) ? (uint32_t) std::stoul(m[1]) : 0u;
        }

        d.desc.abi_version = B200PT_ABI_VERSION;
        d.desc.n_textures = (uint32_t) d.textures.size(); d.desc.textures = d.textures.data();
        d.desc.n_bsdfs = (uint32_t) d.bsdfs.size(); d.desc.bsdfs = d.bsdfs.data();
        d.desc.n_shapes = (uint32_t) d.shapes.size(); d.desc.shapes = d.shapes.data();
        d.desc.n_emitters = (uint32_t) d.emitters.size(); d.desc.emitters = d.emitters.data();
    }

    static double sensor_near(const Sensor *s) {
        auto pc = dynamic_cast<const ProjectiveCamera<Float, Spectrum> *>(s);
        if (!pc) Throw("b200_path_native: perspective sensors only");
        return pc->near_clip();
    }
    static double sensor_far(const Sensor *s) {
        auto pc = dynamic_cast<const ProjectiveCamera<Float, Spectrum> *>(s);
        return pc->far_clip();
    }

    static ScalarTransform4f to_scalar_transform(const Collector::Entry &e) {
        using Transform4f_ = AffineTransform<Point<Float, 4>>;
        if (*e.type == typeid(ScalarTransform4f))
            return *(const ScalarTransform4f *) e.ptr;
        if (*e.type == typeid(Transform4f_)) {
            const Transform4f_ &t = *(const Transform4f_ *) e.ptr;
            ScalarTransform4f out;
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c)
                    out.matrix(r, c) = (float) dr::slice(t.matrix(r, c), 0);
            out.inverse_transpose = dr::transpose(dr::inverse(out.matrix));
            return out;
        }
        Throw("b200_path_native: unexpected type of the sensor's to_world");
    }

    /// Plain dump of the description: "name dtype count\n" + raw little-endian data, one section per array
    static void write_dump(const Description &d, const char *path) {
        FILE *f = std::fopen(path, "wb");
        if (!f) Throw("b200_path_native: cannot write %s", path);
        auto sec = [&](const std::string &name, const char *dtype, const void *p, size_t count, size_t elem) {
            std::fprintf(f, "%s %s %zu\n", name.c_str(), dtype, count);
            std::fwrite(p, elem, count, f);
        };
        uint32_t counts[4] = { d.desc.n_textures, d.desc.n_bsdfs, d.desc.n_shapes, d.desc.n_emitters };
        sec("counts", "u32", counts, 4, 4);
        for (size_t i = 0; i < d.textures.size(); ++i) sec("tex" + std::to_string(i) + ".value", "f32", d.textures[i].value, 3, 4);
        for (size_t i = 0; i < d.bsdfs.size(); ++i) {
            int32_t v[4] = { d.bsdfs[i].type, d.bsdfs[i].twosided, d.bsdfs[i].tex[0], 0 };
            sec("bsdf" + std::to_string(i), "i32", v, 4, 4);
        }
        for (size_t i = 0; i < d.shapes.size(); ++i) {
            const b200pt_shape &s = d.shapes[i];
            int32_t v[6] = { (int32_t) s.n_vertices, (int32_t) s.n_faces, (int32_t) s.layout, s.bsdf, s.emitter, s.sampling };
            sec("shape" + std::to_string(i), "i32", v, 6, 4);
            sec("shape" + std::to_string(i) + ".vertices", "f32", s.vertices, (size_t) s.n_vertices * 8, 4);
            sec("shape" + std::to_string(i) + ".faces", "u32", s.faces, (size_t) s.n_faces * 4, 4);
            float rect[20];
            std::memcpy(rect, s.to_world, 16 * sizeof(float)); std::memcpy(rect + 16, s.frame_n, 3 * sizeof(float)); rect[19] = s.inv_area;
            sec("shape" + std::to_string(i) + ".rect", "f32", rect, 20, 4);
        }
        for (size_t i = 0; i < d.emitters.size(); ++i) {
            const b200pt_emitter &e = d.emitters[i];
            int32_t v[3] = { e.shape, e.radiance_tex, e.type };
            sec("emitter" + std::to_string(i), "i32", v, 3, 4);
            sec("emitter" + std::to_string(i) + ".weight", "f32", &e.sampling_weight, 1, 4);
        }
        const b200pt_sensor &se = d.desc.sensor;
        sec("sensor.sample_to_camera", "f32", se.sample_to_camera, 16, 4);
        sec("sensor.to_world", "f32", se.to_world, 16, 4);
        float clips[3] = { se.near_clip, se.far_clip, se.rfilter_stddev };
        sec("sensor.clips_stddev", "f32", clips, 3, 4);
        uint32_t ints[8] = { se.film_size[0], se.film_size[1], se.crop_size[0], se.crop_size[1], se.crop_offset[0], se.crop_offset[1], (uint32_t) se.rfilter, se.base_seed };
        sec("sensor.ints", "u32", ints, 8, 4);
        std::fclose(f);
    }

    int m_max_depth, m_rr_depth, m_device;
    B200Api m_api;
    b200pt_scene *m_handle = nullptr;
    const Scene *m_scene = nullptr;
};

MI_EXPORT_PLUGIN(B200PathNative)
NAMESPACE_END(mitsuba)
