/*
 * b200pt.h -- C ABI of the B200-native wavefront path tracer (libb200pt.so)
 *
 * This is the drop-in boundary for ONE hot path of Mitsuba 3: the `path`
 * integrator loop and its PRB adjoint together with everything they call
 * down to the ray/triangle tests. Each entry point
 * names the reference interface it replaces (paths relative to the
 * reference checkout).
 *
 * Conventions
 *   - plain C, opaque handles, POD descriptors, no C++/torch types;
 *   - every function returns a b200pt_status (0 = ok); the message of the
 *     last failure on the calling thread is b200pt_last_error();
 *   - the caller owns every host buffer it passes; the library copies what
 *     it needs during the call and owns all device memory;
 *   - "host" pointers are ordinary (ideally pinned) CPU memory, "device"
 *     pointers are CUDA device pointers on the scene's device;
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with B200PT_ERR_CUDA.
 */
#ifndef B200PT_H
#define B200PT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PT_ABI_VERSION 2

#if defined(__GNUC__)
#define B200PT_API __attribute__((visibility("default")))
#else
#define B200PT_API
#endif

typedef enum b200pt_status {
    B200PT_OK              = 0,
    B200PT_ERR_INVALID     = 1, /* bad argument / inconsistent descriptor   */
    B200PT_ERR_CUDA        = 2, /* CUDA runtime failure or no device        */
    B200PT_ERR_UNSUPPORTED = 3, /* feature outside the hot-path scope       */
    B200PT_ERR_NOMEM       = 4
} b200pt_status;

/* ------------------------------------------------------------------------
 * Scene description (host side, POD). Mirrors what the plugin extracts from
 * the live Mitsuba objects: packed meshes (mesh_utils.h:19-46), BSDF /
 * emitter parameters as seen by mi.traverse(), sensor + film + rfilter.
 * ---------------------------------------------------------------------- */

/* Texture: constant `rgb`/float value or a raw float32 `bitmap`
 * (src/textures/bitmap.cpp:496-519, drjit/texture_impl.h:87-205). */
enum { B200PT_TEX_CONST = 0, B200PT_TEX_BITMAP = 1, B200PT_TEX_CHECKERBOARD = 2 /* src/textures/checkerboard.cpp:70-110, constant colours */ };
enum { B200PT_WRAP_REPEAT = 0, B200PT_WRAP_MIRROR = 1, B200PT_WRAP_CLAMP = 2 };
enum { B200PT_FILTER_BILINEAR = 0, B200PT_FILTER_NEAREST = 1 };

typedef struct b200pt_texture {
    int32_t kind;           /* B200PT_TEX_*                                 */
    int32_t channels;       /* 1 or 3                                       */
    float   value[3];       /* B200PT_TEX_CONST (channels==1: value[0]); checkerboard color0 */
    float   value1[3];      /* B200PT_TEX_CHECKERBOARD color1 (gradient: color0 then color1) */
    int32_t width, height;  /* B200PT_TEX_BITMAP                            */
    const float *data;      /* host, height*width*channels, row-major       */
    int32_t wrap;           /* B200PT_WRAP_*                                */
    int32_t filter;         /* B200PT_FILTER_*                              */
    float   to_uv[9];       /* row-major 3x3 uv transform (identity default)*/
    int32_t differentiable; /* !=0: PRB accumulates a gradient buffer       */
} b200pt_texture;

/* BSDF models on the hot path. */
enum {
    B200PT_BSDF_DIFFUSE    = 0, /* src/bsdfs/diffuse.cpp    */
    B200PT_BSDF_CONDUCTOR  = 1, /* src/bsdfs/conductor.cpp  */
    B200PT_BSDF_DIELECTRIC = 2, /* src/bsdfs/dielectric.cpp */
    B200PT_BSDF_PRINCIPLED = 3, /* src/bsdfs/principled.cpp */
    B200PT_BSDF_PLASTIC    = 4  /* src/bsdfs/plastic.cpp (smooth plastic) */
};

/* Texture slots (indices into b200pt_bsdf::tex). */
enum {
    /* diffuse */
    B200PT_SLOT_REFLECTANCE = 0,
    /* conductor / roughconductor (alpha slots: B200PT_M_ROUGH only) */
    B200PT_SLOT_ETA = 0, B200PT_SLOT_K = 1, B200PT_SLOT_SPEC_REFL = 2,
    B200PT_SLOT_ALPHA_U = 3, B200PT_SLOT_ALPHA_V = 4,
    /* plastic */
    B200PT_SLOT_PL_DIFFUSE = 0, B200PT_SLOT_PL_SPEC_REFL = 1,
    /* dielectric / roughdielectric */
    B200PT_SLOT_D_SPEC_REFL = 0, B200PT_SLOT_D_SPEC_TRANS = 1,
    B200PT_SLOT_D_ALPHA_U = 2, B200PT_SLOT_D_ALPHA_V = 3,
    /* principled (principled.cpp:190-330) */
    B200PT_SLOT_P_BASE_COLOR = 0, B200PT_SLOT_P_ROUGHNESS = 1,
    B200PT_SLOT_P_ANISOTROPIC = 2, B200PT_SLOT_P_METALLIC = 3,
    B200PT_SLOT_P_SPEC_TRANS = 4, B200PT_SLOT_P_SPECULAR = 5,
    B200PT_SLOT_P_SPEC_TINT = 6, B200PT_SLOT_P_SHEEN = 7,
    B200PT_SLOT_P_SHEEN_TINT = 8, B200PT_SLOT_P_FLATNESS = 9,
    B200PT_SLOT_P_CLEARCOAT = 10, B200PT_SLOT_P_CLEARCOAT_GLOSS = 11,
    B200PT_MAX_SLOTS = 12
};

/* Principled feature mask == the m_has_* booleans (principledhelpers.h). */
enum {
    B200PT_P_HAS_CLEARCOAT = 1u << 0, B200PT_P_HAS_SHEEN = 1u << 1,
    B200PT_P_HAS_SPEC_TRANS = 1u << 2, B200PT_P_HAS_METALLIC = 1u << 3,
    B200PT_P_HAS_SPEC_TINT = 1u << 4, B200PT_P_HAS_SHEEN_TINT = 1u << 5,
    B200PT_P_HAS_ANISOTROPIC = 1u << 6, B200PT_P_HAS_FLATNESS = 1u << 7,
    B200PT_P_ETA_SPECULAR = 1u << 8  /* eta given explicitly, not `specular` */
};

/* Microfacet variants of B200PT_BSDF_CONDUCTOR / _DIELECTRIC: `roughconductor`
 * (src/bsdfs/roughconductor.cpp) and `roughdielectric` (src/bsdfs/roughdielectric.cpp) are the
 * same types with B200PT_M_ROUGH set and the alpha_u / alpha_v slots filled; the
 * distribution is Beckmann unless B200PT_M_GGX (microfacet.h:36-43), visible-normal
 * sampling only (`sample_visible = true`, the default). */
enum { B200PT_M_ROUGH = 1u << 16, B200PT_M_GGX = 1u << 17,
       B200PT_M_NONLINEAR = 1u << 18 /* plastic: `nonlinear` (plastic.cpp:176) */ };

typedef struct b200pt_bsdf {
    int32_t  type;                  /* B200PT_BSDF_*                        */
    int32_t  twosided;              /* wrapped in `twosided` (twosided.cpp) */
    int32_t  tex[B200PT_MAX_SLOTS]; /* texture index per slot, -1 = unset   */
    float    eta;                   /* dielectric/principled: int_ior/ext_ior*/
    float    spec_srate;            /* principled: main_specular_sampling_rate */
    float    clearcoat_srate;       /* principled: clearcoat_sampling_rate  */
    float    diff_refl_srate;       /* principled: diffuse_reflectance_sampling_rate */
    uint32_t flags;                 /* B200PT_P_* | B200PT_M_*              */
    float    plastic_fdr_int;       /* plastic: fresnel_diffuse_reflectance(1/eta) (plastic.cpp:199, fresnel.h:326-360) */
    float    plastic_spec_weight;   /* plastic: m_specular_sampling_weight (plastic.cpp:202-208)                       */
} b200pt_bsdf;

/* How an emitter's shape is sampled by position
 * (rectangle.cpp:159-179, mesh.cpp:1662-1712). */
enum { B200PT_SAMPLING_NONE = 0, B200PT_SAMPLING_RECTANGLE = 1, B200PT_SAMPLING_MESH = 2 };

/* Layout bits of the packed records (mesh_utils.h:40-46). */
enum { B200PT_LAYOUT_NORMALS = 1, B200PT_LAYOUT_TANGENTS = 2, B200PT_LAYOUT_TEXCOORDS = 4 };

typedef struct b200pt_shape {
    uint32_t n_vertices, n_faces;
    const float    *vertices; /* n_vertices*8: pos3, normal3 (or, with B200PT_LAYOUT_TANGENTS, the three floats of frame_encode(normal, tangent), mesh_utils.h:74-118), uv2 */
    const uint32_t *faces;    /* n_faces*4: v0, v1, v2, flags (bit 31 = FaceUVFlipped, mesh_utils.h:32; read with B200PT_LAYOUT_TANGENTS) */
    uint32_t layout;          /* B200PT_LAYOUT_*                             */
    int32_t  bsdf;            /* index into b200pt_scene_desc::bsdfs         */
    int32_t  emitter;         /* index into ::emitters, -1 = not emissive    */
    int32_t  sampling;        /* B200PT_SAMPLING_*                           */
    float    to_world[16];    /* row-major; B200PT_SAMPLING_RECTANGLE only   */
    float    frame_n[3];      /* rectangle m_frame.n                         */
    float    inv_area;        /* rectangle m_inv_surface_area                */
} b200pt_shape;

/* Emitters: area light (src/emitters/area.cpp:83-209) attached to a shape, or ONE
 * environment emitter per scene (scene.cpp:63-67): `constant`
 * (src/emitters/constant.cpp:60-160) or `envmap` (src/emitters/envmap.cpp:107-590).
 * The order of ::emitters is the order of Scene::emitters() (scene.cpp:45-61): it
 * decides which emitter a sample picks (scene.cpp:248-271). */
enum { B200PT_EMITTER_AREA = 0, B200PT_EMITTER_CONSTANT = 1, B200PT_EMITTER_ENVMAP = 2 };

typedef struct b200pt_emitter {
    int32_t shape;        /* area: index into ::shapes; -1 for environment emitters  */
    int32_t radiance_tex; /* area / constant: texture index (constant: kind CONST)   */
    float   sampling_weight;
    int32_t type;         /* B200PT_EMITTER_*                                        */
    /* envmap: `radiance_tex` is a B200PT_TEX_BITMAP texture (3 channels) holding the
     * latitude-longitude map, row-major height x width, REAL columns only (the `data`
     * parameter of the plugin without its two halo columns, envmap.cpp:155-192); at least
     * 2 x 3 texels (Bitmap::pad_to, envmap.cpp:141). Its wrap / filter fields are ignored.
     * The library adds the periodic halo, and builds the luminance x sin(theta)
     * Hierarchical2D warp (envmap.cpp:474-529, core/distr_2d.h:403-540) and the bounding
     * sphere of the scene (envmap.cpp:260-274) itself; b200pt_scene_update_texture on that
     * texture rebuilds them (EnvironmentMapEmitter::parameters_changed, envmap.cpp:207-258).
     * If the texture is differentiable, PRB accumulates d/d(data) into its gradient. */
    float   env_scale;                /* `scale` (envmap.cpp:196)                    */
    int32_t env_mis_compensation;     /* `mis_compensation` (envmap.cpp:197,497-517) */
    float   to_world[16];             /* row-major; only the linear 3x3 part is used */
    float   to_world_inv[16];         /* row-major inverse of to_world               */
} b200pt_emitter;

/* Reconstruction filter of the film (imageblock.cpp:192-574). */
enum {
    B200PT_RFILTER_BOX = 0,
    B200PT_RFILTER_GAUSSIAN = 1,      /* analytic, polynomial form (gaussian.cpp:57-96; llvm/scalar) */
    B200PT_RFILTER_GAUSSIAN_EXP2 = 2, /* analytic, exp2 form (gaussian.cpp:98-101; cuda)             */
    B200PT_RFILTER_GAUSSIAN_TABLE = 3 /* 31-bin table (rfilter.h:70-79; scalar variants only)        */
};

/* Perspective sensor + hdrfilm (perspective.cpp:239-279, sensor.h:234-269). */
typedef struct b200pt_sensor {
    float    sample_to_camera[16]; /* row-major, inverse perspective_projection */
    float    to_world[16];         /* row-major camera-to-world                 */
    float    near_clip, far_clip;
    uint32_t film_size[2];         /* width, height                             */
    uint32_t crop_size[2];
    uint32_t crop_offset[2];
    int32_t  rfilter;              /* B200PT_RFILTER_*                          */
    float    rfilter_stddev;       /* gaussian (default 0.5; radius = 4 stddev) */
    uint32_t base_seed;            /* sampler `seed` property (sampler.cpp:133) */
} b200pt_sensor;

typedef struct b200pt_scene_desc {
    uint32_t abi_version; /* B200PT_ABI_VERSION */
    uint32_t n_shapes;   const b200pt_shape   *shapes;
    uint32_t n_bsdfs;    const b200pt_bsdf    *bsdfs;
    uint32_t n_emitters; const b200pt_emitter *emitters;
    uint32_t n_textures; const b200pt_texture *textures;
    b200pt_sensor sensor;
} b200pt_scene_desc;

/* Integrator properties (integrator.cpp:26-29,130-146,539-550) plus the
 * pixel-tile shard this process renders. */
typedef struct b200pt_render_params {
    uint32_t seed;          /* `seed` argument of Integrator::render          */
    uint32_t spp;           /* samples per pixel                              */
    int32_t  max_depth;     /* -1 = unbounded                                 */
    int32_t  rr_depth;      /* default 5                                      */
    int32_t  hide_emitters;
    uint32_t shard_rank;    /* this process' rank in [0, shard_count)         */
    uint32_t shard_count;   /* 1 = whole frame                                */
    uint32_t tile_size;     /* pixel-tile edge for sharding (default 32); tile (tx, ty) belongs to
                               rank (tx + ty * (shard_count / 2 + 1)) % shard_count            */
    uint32_t chunk_lanes;   /* wavefront lanes per pass, 0 = library default  */
    int32_t  prb;           /* 0 = `path` estimator, 1 = `prb` primal rules   */
} b200pt_render_params;

/* Counters of the last render call on a scene (measurement). */
typedef struct b200pt_stats {
    uint64_t samples;          /* camera samples traced                      */
    uint64_t bounces;          /* loop iterations over all samples           */
    uint64_t shadow_rays;      /* any-hit queries issued                     */
    uint64_t kernel_launches;  /* kernels of this library launched           */
    double   device_ms;        /* CUDA-event time of the whole call          */
    double   trace_ms;         /* CUDA-event time inside the closest-hit kernel */
    uint64_t trace_launches;
    uint64_t trace_rays;       /* closest-hit rays processed                 */
} b200pt_stats;

typedef struct b200pt_scene b200pt_scene;

/* ---- library ---------------------------------------------------------- */
B200PT_API uint32_t    b200pt_abi_version(void);
B200PT_API const char *b200pt_last_error(void);
/* Number of visible CUDA devices (0 without a driver/GPU). */
B200PT_API int         b200pt_device_count(void);
/* Devices this process renders on. The multi-GPU layout is one process per GPU: `ids` is the list of
 * CUDA ordinals of the job on this node and `b200pt_scene_create(desc, B200PT_DEVICE_AUTO, ..)` picks
 * ids[LOCAL_RANK % n] (LOCAL_RANK from the environment, 0 if unset) -- what `jit_set_device` / `cuda_ad_rgb`'s device
 * selection is to the reference (drjit-core jit.h: jit_cuda_set_device). n = 0 restores the default (device 0). */
#define B200PT_DEVICE_AUTO (-1)
B200PT_API b200pt_status b200pt_set_devices(int n, const int *ids);

/* ---- scene life cycle: replaces Scene::Scene accel build (scene.cpp:93,
 *      scene_optix.inl:446) for triangle meshes ------------------------- */
B200PT_API b200pt_status b200pt_scene_create(const b200pt_scene_desc *desc, int device,
                                  b200pt_scene **out);
B200PT_API void          b200pt_scene_destroy(b200pt_scene *scene);
/* Geometry update with unchanged topology (params['<mesh>.vertex_positions'] = ...; params.update() ->
 * Mesh::parameters_changed + Scene::parameters_changed -> accel update, mesh.cpp:170-230, scene.cpp:517-540): replace
 * the packed (V, 8) vertex records of shape `shape` and REFIT the BVH on the device (boxes recomputed bottom-up, tree
 * topology and triangle order kept; a large deformation costs traversal efficiency, never correctness). Shapes that are
 * sampled as emitters carry host-built sampling tables: for those the call fails with B200PT_ERR_UNSUPPORTED and the
 * scene has to be created again. */
B200PT_API b200pt_status b200pt_scene_update_vertices(b200pt_scene *scene, uint32_t shape, const float *vertices,
                                           uint32_t n_vertices);
/* Parameter update after an optimiser step (SceneParameters.update ->
 * parameters_changed, util.py:272-338): overwrite texture `tex` with `n`
 * floats (3/1 for constants, h*w*c for bitmaps). */
B200PT_API b200pt_status b200pt_scene_update_texture(b200pt_scene *scene, uint32_t tex,
                                          const float *host_data, size_t n);

/* ---- forward render: replaces SamplingIntegrator::render (JIT branch,
 *      integrator.cpp:275-389) + PathIntegrator::sample (path.cpp:94-346)
 *      + ImageBlock::put + HDRFilm::develop ------------------------------ */
/* Host entry point (what the plugin calls): out_host is H*W*3 floats. */
B200PT_API b200pt_status b200pt_render(b200pt_scene *scene, const b200pt_render_params *p,
                            float *out_host);
/* Device entry points for multi-GPU: accumulate this shard's samples into a
 * raw film block (H*W*4: R,G,B,weight) owned by the caller, then develop
 * (hdrfilm.cpp:393) after the caller has all-reduced the block. All work is
 * enqueued on `cuda_stream` (a cudaStream_t; NULL = the CUDA default stream) so
 * that it orders with the caller's collectives; the call returns after the
 * stream has drained (statistics are read back). */
B200PT_API b200pt_status b200pt_render_accumulate(b200pt_scene *scene,
                                       const b200pt_render_params *p,
                                       float *film_device /* H*W*4, zeroed by caller */,
                                       void *cuda_stream);
B200PT_API b200pt_status b200pt_develop(b200pt_scene *scene, const float *film_device,
                             float *out_device /* H*W*3 */, void *cuda_stream);

/* ---- adjoint: replaces RBIntegrator.render_backward (common.py:625-783)
 *      + PRBIntegrator.sample Backward mode (prb.py:68-339) -------------- */
/* grad_in_host: H*W*3 (dLoss/dImage). Gradients are ACCUMULATED into the
 * per-texture gradient buffers of the scene (dr.grad accumulation). */
B200PT_API b200pt_status b200pt_render_backward(b200pt_scene *scene,
                                     const b200pt_render_params *p,
                                     const float *grad_in_host);
B200PT_API b200pt_status b200pt_render_backward_device(b200pt_scene *scene,
                                            const b200pt_render_params *p,
                                            const float *grad_in_device,
                                            void *cuda_stream);
/* ---- forward mode: replaces RBIntegrator.render_forward (common.py:560-623) + PRBIntegrator.sample
 *      Forward mode. The parameter tangents (dr.set_grad of the reference) are written per
 *      differentiable texture, in the layout of the texture's data; out_host: H*W*3 = d(image). */
B200PT_API b200pt_status b200pt_tangent_zero(b200pt_scene *scene);
B200PT_API b200pt_status b200pt_tangent_write(b200pt_scene *scene, uint32_t tex,
                                   const float *host_in, size_t n);
B200PT_API b200pt_status b200pt_render_forward(b200pt_scene *scene,
                                    const b200pt_render_params *p, float *out_host);

B200PT_API b200pt_status b200pt_grad_zero(b200pt_scene *scene);
/* Copy the gradient of differentiable texture `tex` to the host. */
B200PT_API b200pt_status b200pt_grad_read(b200pt_scene *scene, uint32_t tex,
                               float *host_out, size_t n);
/* Device view of all gradient buffers as one flat fp32 array (for one fused
 * NCCL all-reduce); offsets per texture via b200pt_grad_offset. */
B200PT_API b200pt_status b200pt_grad_device_view(b200pt_scene *scene, float **ptr, size_t *n);
B200PT_API b200pt_status b200pt_grad_offset(b200pt_scene *scene, uint32_t tex,
                                 size_t *offset, size_t *n);

/* ---- operators on the path, exposed for parity tests ------------------ */
/* Scene::ray_intersect_preliminary (scene.cpp:216): n rays, rays_host =
 * n*7 floats (o3, d3, maxt). Outputs: t (inf on miss), prim_uv n*2,
 * prim_index (within shape), shape_index (-1 on miss). */
B200PT_API b200pt_status b200pt_ray_intersect(b200pt_scene *scene, uint32_t n,
                                   const float *rays_host, float *t_out,
                                   float *uv_out, uint32_t *prim_out,
                                   int32_t *shape_out);
/* Scene::ray_test (scene.cpp:232): hit_out[i] = 1 if occluded. */
B200PT_API b200pt_status b200pt_ray_test(b200pt_scene *scene, uint32_t n,
                              const float *rays_host, uint8_t *hit_out);
/* BSDF::eval_pdf_sample (bsdf.cpp:21-31) for n queries on BSDF `bsdf`:
 * in_host = n*10 floats (wi3, wo3, uv2 ... see layout below),
 *   [0..2] si.wi (local), [3..5] wo (local), [6..7] si.uv,
 *   [8] sample1, [9..10] sample2  -> stride 11
 * out_host = n*14 floats:
 *   [0..2] eval (f*cos), [3] pdf, [4..6] bs.wo, [7] bs.pdf, [8] bs.eta,
 *   [9] sampled_type (as float bits of uint32), [10..12] weight, [13] sampled_component */
B200PT_API b200pt_status b200pt_bsdf_eval_pdf_sample(b200pt_scene *scene, uint32_t bsdf,
                                          uint32_t n, const float *in_host,
                                          float *out_host);

/* Emitter::sample_direction / eval / pdf_direction of the scene's environment emitter
 * (envmap.cpp:276-396, constant.cpp:95-152; Python: src/render/python/emitter_v.cpp),
 * WITHOUT the scene's emitter-selection pmf. in_host = n*8 floats:
 *   [0..2] reference point it.p, [3..4] sample, [5..7] a query direction d_q (unit, world)
 * out_host = n*20 floats:
 *   [0..2] ds.d, [3] ds.pdf, [4] ds.dist, [5..6] ds.uv, [7..9] weight (radiance / pdf),
 *   [10..12] eval(-wi = ds.d), [13] pdf_direction(ds.d),
 *   [14..16] eval(-wi = d_q), [17] pdf_direction(d_q), [18..19] unused */
B200PT_API b200pt_status b200pt_env_query(b200pt_scene *scene, uint32_t n, const float *in_host,
                               float *out_host);

/* ---- measurement ------------------------------------------------------ */
B200PT_API b200pt_status b200pt_get_stats(b200pt_scene *scene, b200pt_stats *out);
/* sizeof() of the ABI structs as compiled: 0 texture, 1 bsdf, 2 shape, 3 emitter,
 * 4 sensor, 5 scene_desc, 6 render_params, 7 stats (binding self-check). */
B200PT_API size_t b200pt_abi_sizeof(int which);

#ifdef __cplusplus
}
#endif
#endif /* B200PT_H */
