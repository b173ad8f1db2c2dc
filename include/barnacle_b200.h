/* barnacle_b200.h — C ABI of the B200 path-tracing hot path.
 *
 * This is the drop-in boundary for Barnacle's render loop
 * (reference: IntegratorBase.Render, Base/Integrator.fs:9-11, called once by
 * Scene.Render, Extensions/Scene/Render.fs:16).  A managed host (F#,
 * `DllImport`) flattens its object graph into the POD arrays described by
 * BnSceneDesc and calls bn_scene_create / bn_render.  All structs are
 * little-endian POD; every pointer is a plain host pointer owned by the
 * caller; nothing here mentions torch, CUDA runtime or C++ types.
 *
 * Layouts that already exist in the reference are kept byte-for-byte so the
 * managed side can pin its own arrays:
 *   BnBVHNode    = BVHNode, 32 B explicit layout   (Util/BVH.fs:52-73)
 *   BnAliasEntry = AliasTable.Entry {alias,prob,pdf} (Util/AliasTable.fs:7-12)
 *   float[3]     = System.Numerics.Vector3 packed (12-B stride)
 *   float[16]    = Matrix4x4 row-major M11..M44, ROW-VECTOR convention
 *                  (v' = v*M, translation in M41..M43)
 *   int32[3]     = TriangleIndex                   (Mesh.fs:113-117)
 *   film         = Film.Pixels, Vector3[W*H], Y-flipped (Base/Film.fs:17,41-46)
 */
#ifndef BARNACLE_B200_H
#define BARNACLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BN_API __attribute__((visibility("default")))

/* ---- status codes (reference convention is `failwith`; the wrapper does
 *      `if rc <> 0 then failwith (bn_last_error())`) ------------------------ */
enum {
  BN_OK = 0,
  BN_ERR_INVALID = -1,   /* bad argument / inconsistent scene description */
  BN_ERR_CUDA = -2,      /* CUDA runtime error (sticky: scene unusable)   */
  BN_ERR_NO_DEVICE = -3, /* no usable sm_100 device: there is NO CPU fallback */
  BN_ERR_IO = -4,        /* host-side loader: file/JSON/OBJ problem       */
  BN_ERR_NO_LIGHT = -5   /* LightSamplerBase: "No light primitives found" (Base/LightSampler.fs:8-9) */
};

/* ---- blittable records ---------------------------------------------------- */

/* Util/BVH.fs:52-73.  Preorder array, left child = i+1. */
typedef struct BnBVHNode {
  float bounds_min[3];        /*  0 */
  float bounds_max[3];        /* 12 */
  int32_t right_or_offset;    /* 24  RightChild (interior) | InstanceOffset (leaf) */
  uint8_t is_leaf;            /* 28 */
  int8_t split_axis;          /* 29 */
  int8_t count;               /* 30  instanceCount (leaf) */
  uint8_t visibility_mask;    /* 31  unused by the reference */
} BnBVHNode;

/* Util/AliasTable.fs:7-12 */
typedef struct BnAliasEntry {
  int32_t alias;
  float prob;
  float pdf;
} BnAliasEntry;

enum { BN_PRIM_MESH = 0, BN_PRIM_SPHERE = 1 };

/* One PrimitiveInstance (Base/Primitive.fs:100-141), in TLAS order (i.e. after
 * BVHNode.Build permuted the instance array, Util/BVH.fs:244-246). */
typedef struct BnInstance {
  uint32_t prim_kind;         /* BN_PRIM_MESH | BN_PRIM_SPHERE */
  uint32_t prim_id;           /* index into meshes[] or sphere_radii[] */
  int32_t material_id;        /* -1 = no material (HasMaterial = false) */
  int32_t light_id;           /* -1 = no light    (HasLight = false) */
  float object_to_world[16];
  float world_to_object[16];
  float bounds_min[3];        /* world AABB, AxisAlignedBoundingBox.Transform (Util/BVH.fs:29-40) */
  float bounds_max[3];
} BnInstance;                 /* 168 B */

/* One MeshPrimitive (Extensions/Primitive/Mesh.fs:119-186): slices into the
 * shared vertex / triangle-index / BLAS-node / alias arrays.  Triangle indices
 * are in BLAS order (post BVHNode.Build permutation) and are LOCAL to the
 * mesh's vertex slice. */
typedef struct BnMesh {
  uint32_t vertex_offset, vertex_count;
  uint32_t tri_offset, tri_count;
  uint32_t node_offset, node_count;
  uint32_t alias_offset;      /* tri_count entries */
  uint32_t reserved;
} BnMesh;

enum { BN_MAT_LAMBERTIAN = 0, BN_MAT_MIRROR = 1, BN_MAT_DIELECTRIC = 2, BN_MAT_PBR = 3 };

/* Lambertian.fs:10 / Mirror.fs:10 / Dielectric.fs:11-12 / PBR.fs:10-12 */
typedef struct BnMaterial {
  uint32_t type;
  float base_color[3];
  float p0;                   /* dielectric: IOR | pbr: Metallic (already clamped to [0,1]) */
  float p1;                   /* pbr: Alpha = max(roughness^2, 1e-3) */
} BnMaterial;

/* DiffuseLight (Base/Light.fs:26-53) */
typedef struct BnLight {
  float emission[3];
  uint32_t two_sided;
} BnLight;

enum { BN_CAM_PINHOLE = 0, BN_CAM_THIN_LENS = 1 };

/* Pinhole.fs:7-10, ThinLens.fs:8-11, Camera.fs:6-8 */
typedef struct BnCamera {
  uint32_t type;
  float fov_y;                /* degrees */
  float aspect_ratio;
  float aperture;
  float focus_distance;
  float push_forward;
  float camera_to_world[16];
} BnCamera;

typedef struct BnSceneDesc {
  const BnBVHNode* tlas_nodes;     uint32_t tlas_node_count;
  const BnInstance* instances;     uint32_t instance_count;
  /* indices (into instances[]) of the emissive instances, in TLAS order —
   * LightSamplerBase.Instances (Base/LightSampler.fs:7,10) */
  const uint32_t* light_instances; uint32_t light_instance_count;
  const BnMesh* meshes;            uint32_t mesh_count;
  const float* vertices;           uint32_t vertex_count;    /* xyz packed */
  const int32_t* triangles;        uint32_t triangle_count;  /* i0 i1 i2 packed */
  const BnBVHNode* blas_nodes;     uint32_t blas_node_count;
  const BnAliasEntry* alias;       uint32_t alias_count;
  const float* sphere_radii;       uint32_t sphere_count;
  const BnMaterial* materials;     uint32_t material_count;
  const BnLight* lights;           uint32_t light_count;
  BnCamera camera;
} BnSceneDesc;

/* ProgressiveIntegrator / PathTracingIntegrator parameters
 * (Base/Integrator.fs:14-19, PathTracing.fs:9-12) plus the sharding window. */
typedef struct BnRenderParams {
  int32_t width, height;      /* Film.ImageWidth/Height */
  int32_t spp;                /* SamplePerPixel — also the 1/spp weight */
  int32_t max_depth;          /* default 8  (Loader.fs:182) */
  int32_t rr_depth;           /* default 5  (Loader.fs:183) */
  int32_t frame_id;           /* ProgressiveIntegrator.FrameId */
  /* sharding window: only sampleIds in [sample_begin, sample_end) of pixels in
   * [x0,x1) x [y0,y1) are rendered; other film pixels are written as 0.
   * A full render is sample 0..spp, rect 0,0,width,height. */
  int32_t sample_begin, sample_end;
  int32_t x0, y0, x1, y1;
  uint32_t flags;             /* BN_RENDER_* */
  /* tile-row interleave (multi-GPU tile split): of the 16-pixel tile rows of the
   * window (the reference's tile size, Integrator.fs:16), counted from y0, only
   * rows r with r % interleave_count == interleave_index are rendered.
   * interleave_count <= 1 renders every row. */
  int32_t interleave_count, interleave_index;
  /* which ProgressiveIntegrator.Li runs (Loader.fs:185-188): BN_INTEGRATOR_* ; 0 = path tracing */
  int32_t integrator;
} BnRenderParams;

enum {
  BN_INTEGRATOR_PATH_TRACING = 0, /* PathTracingIntegrator.Li, Extensions/Integrator/PathTracing.fs:14-81 */
  BN_INTEGRATOR_DIRECT = 1,       /* DirectIntegrator.Li,      Extensions/Integrator/Direct.fs:10-40 (max_depth / rr_depth ignored) */
  BN_INTEGRATOR_NORMAL = 2        /* NormalIntegrator.Li,      Extensions/Integrator/Normal.fs:10-17 */
};

enum {
  BN_RENDER_DEFAULT = 0,
  /* trace NEE shadow rays even when the BSDF evaluates to exactly 0 (mirror /
   * dielectric): reference-equivalent ray counts, same image */
  BN_RENDER_TRACE_NULL_SHADOW = 1u << 0,
  /* bracket every kernel launch with CUDA events on the launching stream and
   * report per-class device time in BnStats.{extend,shade,shadow,other}_ms */
  BN_RENDER_PROFILE = 1u << 1,
  /* route EVERY ray through the exact (reference-op-for-op) traversal of the
   * fix-up kernel instead of the fast phases: slow; exists so tests can show the
   * two paths agree bit for bit on real path-tracing rays */
  BN_RENDER_FORCE_EXACT = 1u << 2
};

typedef struct BnStats {
  uint64_t paths;             /* camera paths started */
  uint64_t extend_rays;       /* closest-hit rays traced */
  uint64_t shadow_rays;       /* any-hit rays traced */
  uint64_t shadow_rays_ref;   /* any-hit rays the reference would trace (Q5) */
  uint64_t kernel_launches;   /* this library's kernels launched by the call */
  double gpu_ms;              /* device time of the render (CUDA events) */
  double extend_ms, shade_ms, shadow_ms, other_ms; /* filled when profiling flag set */
} BnStats;

typedef struct BnRay {
  float origin[3];
  float direction[3];
  float tmax;                 /* closest hit: initial t (use +inf); any hit: tmax */
} BnRay;                      /* 28 B */

typedef struct BnHit {
  float t;                    /* closest: hit distance, +inf... unchanged tmax on miss; any: 0 */
  float u, v;                 /* LocalGeometry.uv */
  int32_t instance;           /* TLAS-order instance index, -1 on miss; any-hit: 1/0 in `instance` */
  int32_t primitive;          /* BLAS-order triangle index (geom.tag BEFORE Primitive.fs:57 resets it), 0 for spheres */
} BnHit;                      /* 20 B */

typedef struct BnScene BnScene;

/* ---- device library ------------------------------------------------------- */

/* Number of usable CUDA devices (0 if none; never throws). */
BN_API int bn_device_count(void);

/* Measurement aid (bench.py roofline): read bandwidth of an L2-resident buffer of `bytes`
 * (16-B loads that bypass L1, `iters` sweeps), in GB/s. */
BN_API int bn_measure_l2_read_gbs(int device, uint64_t bytes, int iters, double* gbs);

/* bn_scene_destroy parks the scene's device buffers (wave queues: up to ~16 GB at the default 64 Mi-path wave, counters,
 * film, the scene arena) per device for the next scene, so that a host which re-creates the scene every frame pays no
 * driver allocation.  This call returns that memory to the driver (device < 0: every device).  Setting the environment
 * variable BN_NO_BUFFER_CACHE disables the parking altogether. */
BN_API int bn_release_cached_buffers(int device);

/* Thread-local description of the last error on this thread. */
BN_API const char* bn_last_error(void);

/* Copies the scene to `device` (ordinal) and builds the traversal layout.
 * Keeps no host pointers. */
BN_API int bn_scene_create(const BnSceneDesc* desc, int device, BnScene** out);
BN_API void bn_scene_destroy(BnScene* scene);

/* Replaces PathTracingIntegrator's Render (Base/Integrator.fs:46-55 +
 * PathTracing.fs:14-81): renders into film_rgb (HOST pointer, W*H*3 floats,
 * Film.Pixels layout), blocking. `stats` may be NULL. */
BN_API int bn_render(BnScene* scene, const BnRenderParams* params, float* film_rgb, BnStats* stats);

/* Same, but the film is a DEVICE pointer on the scene's device and the work is
 * enqueued on `cuda_stream` (a cudaStream_t passed as void*; NULL = default
 * stream) and synchronised before returning.  This is what the multi-GPU
 * driver uses so the per-GPU films can be handed to NCCL without a host hop. */
BN_API int bn_render_device(BnScene* scene, const BnRenderParams* params, void* d_film_rgb,
                            void* cuda_stream, BnStats* stats);

/* ---- several devices behind one call ---------------------------------------------------------
 * In the reference the whole machine's parallelism lives INSIDE Integrator.Render (Base/Integrator.fs:46-55: 16x16
 * tiles over the TPL pool, called once from Extensions/Scene/Render.fs:15-17), so the drop-in fans out over the GPUs of
 * the box from inside the call as well.  bn_multi_scene_create flattens the scene once and uploads it to every listed
 * device (one host thread each; the same ordinal may be listed twice); bn_render_multi renders each device's share of
 * the (pixel, sampleId) space and combines the films on devices[0] with one kernel that reads the other devices' films
 * over NVLink (peer access; a staged peer copy where there is none), adding them in the order of the list — the same
 * film every run.  No NCCL / torch involved.  film_rgb: HOST pointer, W*H*3 floats, Film.Pixels layout. */
typedef struct BnMultiScene BnMultiScene;
enum {
  BN_PARTITION_AUTO = 0,   /* sample split when the window has at least as many samples as devices, else tile split */
  BN_PARTITION_SAMPLE = 1, /* device g renders sampleIds [b + g*n/G, b + (g+1)*n/G) of every pixel, weight 1/spp: the same sample set as
                              one device; the per-pixel sum is re-associated (fp32 rounding only) */
  BN_PARTITION_TILE = 2    /* device g renders the 16-pixel tile rows g, g+G, ... (Integrator.fs:16 tile size): disjoint pixels, bit-identical
                              to one device */
};
BN_API int bn_multi_scene_create(const BnSceneDesc* desc, const int32_t* devices, int32_t n_devices, BnMultiScene** out);
BN_API void bn_multi_scene_destroy(BnMultiScene* scene);
BN_API int bn_multi_scene_device_count(const BnMultiScene* scene);
/* stats (may be NULL): rays / paths / launches summed over the devices; gpu_ms = the slowest device + the combine. */
BN_API int bn_render_multi(BnMultiScene* scene, const BnRenderParams* params, int32_t partition, float* film_rgb, BnStats* stats);
/* The share device `rank` of `n_devices` renders (host logic only, no device needed): *out = params restricted to it,
 * *empty = 1 when the share holds no path. */
BN_API int bn_multi_partition(const BnRenderParams* params, int32_t partition, int32_t n_devices, int32_t rank, BnRenderParams* out, int32_t* empty);

/* Fixed-batch traversal entry (parity tests, microbenchmarks).
 * any_hit = 0: PrimitiveAggregate.Intersect/3 (Extensions/Aggregate/BVH.fs:37-58)
 * any_hit = 1: PrimitiveAggregate.Intersect/2 (Extensions/Aggregate/BVH.fs:11-35)
 * rays/hits are HOST pointers. */
BN_API int bn_trace(BnScene* scene, const BnRay* rays, uint64_t n, int any_hit, BnHit* hits);

/* Device-resident variant: d_rays / d_hits are device pointers; returns the
 * kernel's device time in *ms (may be NULL). */
BN_API int bn_trace_device(BnScene* scene, const void* d_rays, uint64_t n, int any_hit,
                           void* d_hits, void* cuda_stream, float* ms);

/* Film.PostProcess + Rgba32 conversion (Base/Film.fs:21-30,55-66) on the device: d_film_rgb
 * (W*H*3 float, device) -> d_rgba8 (W*H*4 bytes, device).  tone_mapping: 0 identity, 1 aces, 2 gamma. */
BN_API int bn_film_to_rgba8_device(BnScene* scene, const void* d_film_rgb, int32_t width, int32_t height, int32_t tone_mapping,
                                   void* d_rgba8, void* cuda_stream);

/* Per-path radiance dump for parity tests: Li * 1/pdf for every (pixel, sample)
 * in the window, laid out [sample - sample_begin][(y - y0)*(x1-x0) + (x - x0)][3],
 * HOST pointer. */
BN_API int bn_render_radiance(BnScene* scene, const BnRenderParams* params, float* radiance);

/* ---- BVH build on the device ("next" row N2: BVHNode.Build, Util/BVH.fs:109-247) ----------
 * Same contract as bn_host_bvh_build below (n boxes of 6 floats; nodes in the reference's
 * preorder layout; perm[i] = original index now at slot i; returns the node count, < 0 on
 * error) and the same bytes: node array and permutation are identical to the host builder's.
 * Boxes must be finite.  `ms` (may be NULL) receives the device time of the build. */
BN_API int bn_bvh_build(int device, const float* boxes, uint32_t n, BnBVHNode* nodes, uint32_t max_nodes, uint32_t* perm, float* ms);
/* Device-pointer variant (d_nodes must hold 2n-1 nodes at most; pass max_nodes accordingly). */
BN_API int bn_bvh_build_device(int device, const void* d_boxes, uint32_t n, void* d_nodes, uint32_t max_nodes, void* d_perm,
                               void* cuda_stream, float* ms);

/* ---- PSSMLT ("next" row N1: PSSMLTIntegrator, Extensions/Integrator/PSSMLT.fs) ------------- */

enum { BN_MLT_GAUSSIAN = 0, BN_MLT_KELEMEN = 1 };  /* MutationStrategy, PSSMLT.fs:15-18 */

typedef struct BnMltParams {
  int32_t width, height;
  int32_t mutations_per_pixel; /* PSSMLTIntegrator's `mutationPerPixel` (the JSON `spp`) */
  int32_t max_depth, rr_depth, frame_id;
  int32_t n_bootstrap;         /* default 4*1024*1024 (Loader.fs:190) */
  int32_t n_chains;            /* default 1024        (Loader.fs:191) */
  int32_t strategy;            /* BN_MLT_GAUSSIAN (sigma = p0, default 1e-2) | BN_MLT_KELEMEN (epsMin = p0, epsMax = p1; 1/1024, 1/16) */
  float p0, p1;
  float large_step_prob;       /* default 0.5 (Loader.fs:202) */
  /* multi-GPU shard: only chains [chain_begin, chain_end) are run (the bootstrap is always complete) */
  int32_t chain_begin, chain_end;
} BnMltParams;

typedef struct BnMltStats {
  float b;                     /* PSSMLTIntegrator.B: mean bootstrap luminance (PSSMLT.fs:394) */
  uint32_t reserved;
  uint64_t accepted, proposed; /* AcceptedMutationCount / ProposedMutationCount of the chains run */
  uint64_t rays;               /* closest-hit + any-hit rays traced (bootstrap + chains) */
  double bootstrap_ms, chains_ms; /* device time of the two phases */
} BnMltStats;

/* Replaces PSSMLTIntegrator.Render (PSSMLT.fs:379-414): bootstrap, B, chains, film splats
 * (atomic adds instead of the reference's racy Film.Accumulate).  film_rgb: HOST
 * pointer, W*H*3 floats, Film.Pixels layout; accumulated INTO (the caller clears it, as
 * Scene.Render does). */
BN_API int bn_render_pssmlt(BnScene* scene, const BnMltParams* params, float* film_rgb, BnMltStats* stats);
/* Same with the film on the device (multi-GPU reduce). */
BN_API int bn_render_pssmlt_device(BnScene* scene, const BnMltParams* params, void* d_film_rgb, void* cuda_stream, BnMltStats* stats);
/* Phase 1 only: BootstrapWeights[n_bootstrap] (PSSMLT.fs:247-273), HOST pointer — parity tests. */
BN_API int bn_pssmlt_bootstrap(BnScene* scene, const BnMltParams* params, float* weights);
/* Parity-test entry: bn_render_pssmlt that also returns the accepted-mutation count of every chain it ran
 * (per_chain_accepted[chain_end - chain_begin], HOST pointer) — what PSSMLT.fs:412 sums into AcceptedMutationCount. */
BN_API int bn_debug_render_pssmlt_chains(BnScene* scene, const BnMltParams* params, float* film_rgb, BnMltStats* stats, unsigned int* per_chain_accepted);

/* Parity-test entry for the ordering of the live paths between bounces (no counterpart in the reference: csrc/cuda/ray_sort.cuh):
 * runs the library's own counting sort (histogram, scan, rank) on n 12-bit keys and returns perm[ordered slot] = index
 * (HOST pointers).  A correct result is a permutation of 0..n-1 along which the keys never decrease. */
BN_API int bn_debug_order_keys(int device, const uint16_t* keys, uint32_t n, uint32_t* perm);

/* ---- host-side scene builder (stands in for the managed host: JSON schema of
 *      Extensions/Scene/Loader.fs, Scene.Traverse, BVHNode.Build, AliasTable) -- */

typedef struct BnHostScene BnHostScene;

typedef struct BnHostSceneInfo {
  int32_t width, height;          /* film */
  int32_t tone_mapping;           /* 0 identity, 1 aces, 2 gamma (Base/Film.fs:10-13) */
  int32_t integrator;             /* 0 normal, 1 direct, 2 path-tracing, 3 pssmlt (Loader.fs:185-204) */
  int32_t spp, max_depth, rr_depth;
  /* pssmlt only (Loader.fs:189-203) */
  int32_t n_bootstrap, n_chains;
  int32_t mutation_strategy;      /* BN_MLT_* */
  float large_step_prob;
} BnHostSceneInfo;

/* Scene.Load (Loader.fs:277-281) + Scene.Traverse(t=time) + BVHAggregate +
 * UniformLightSampler construction (Render.fs:11-14).  `base_dir` resolves
 * relative mesh `uri`s (NULL = process CWD, like the reference). */
BN_API int bn_host_scene_load(const char* json_path, const char* base_dir, float time, BnHostScene** out);
BN_API int bn_host_scene_load_string(const char* json_text, const char* base_dir, float time, BnHostScene** out);
/* Same as bn_host_scene_load, with every BVHNode.Build (each mesh's BLAS, Mesh.fs:169-171, and the
 * TLAS, Aggregate/BVH.fs:9) done by bn_bvh_build on CUDA device `build_device` (-1: on the host).
 * The resulting BnSceneDesc is identical either way. */
BN_API int bn_host_scene_load_ex(const char* json_path, const char* base_dir, float time, int build_device, BnHostScene** out);
BN_API const BnSceneDesc* bn_host_scene_desc(const BnHostScene* scene);
BN_API void bn_host_scene_info(const BnHostScene* scene, BnHostSceneInfo* info);
/* original (pre-BVH-permutation) index of TLAS-order instance i / BLAS-order triangle i of mesh m */
BN_API const uint32_t* bn_host_scene_instance_permutation(const BnHostScene* scene);
BN_API const uint32_t* bn_host_scene_triangle_permutation(const BnHostScene* scene, uint32_t mesh);
BN_API void bn_host_scene_destroy(BnHostScene* scene);

/* BVHNode.Build (Util/BVH.fs:239-247) on n boxes (min xyz, max xyz packed, 6
 * floats each).  Writes up to max_nodes nodes, the permutation (perm[i] =
 * original index now at slot i) and returns the node count (<0 on error). */
BN_API int bn_host_bvh_build(const float* boxes, uint32_t n, BnBVHNode* nodes, uint32_t max_nodes, uint32_t* perm);

/* AliasTable ctor (Util/AliasTable.fs:14-51). */
BN_API int bn_host_alias_build(const float* weights, uint32_t n, BnAliasEntry* out);

/* Film.PostProcess + Rgba32 conversion (Base/Film.fs:21-30,55-66): film_rgb
 * (W*H*3 float) -> rgba8 (W*H*4). */
BN_API int bn_host_film_to_rgba8(const float* film_rgb, int32_t width, int32_t height, int32_t tone_mapping, uint8_t* rgba8);

#ifdef __cplusplus
}
#endif
#endif /* BARNACLE_B200_H */
