// Device-resident scene layout (HBM; every record 16-B aligned so all fetches
// are LDG.128).  Built once by bn_scene_create from the reference-layout
// BnSceneDesc; see DESIGN.md "Data layout in HBM".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bn {

// Child reference encoding (GNode.left/right, traversal stack entries)
//   bit31 = 0 : interior node, bits 0..29 = ABSOLUTE GNode index (scene-wide array)
//   bit31 = 1 : leaf
//        BLAS: bits 27..29 = triangle count (<= 7), bits 0..26 = ABSOLUTE first triangle
//        TLAS: bits 0..29  = instance slot (every TLAS leaf is ONE instance, see
//              scene_convert.cpp: a reference leaf of k instances becomes a chain
//              of k-1 order-preserving pseudo nodes)
// During traversal TLAS-level refs additionally carry kTlasBit (bit 30) so a pop
// knows which space the ray is in.  GNode.axis == 3 means "always left first".
constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kMaxLeafFirst = 1u << 27;

// Interior node: BOTH children's boxes in the parent (one 64-B fetch per
// interior visit instead of two dependent 32-B fetches of BVHNode).  Topology
// and child order are exactly the reference's (left = preorder i+1, right =
// RightChild), so the visiting order — and with it every exact-t tie-break —
// is unchanged.
struct __align__(16) GNode {
  float lmin[3], lmax[3];  // left child bounds  (BVHNode.Bounds of node i+1)
  float rmin[3], rmax[3];  // right child bounds (BVHNode.Bounds of RightChild)
  uint32_t left, right;    // child refs
  uint32_t axis;           // BVHNode.SplitAxis
  uint32_t pad;
};
static_assert(sizeof(GNode) == 64, "GNode must be 64 B");

struct __align__(16) GTri {  // pre-gathered vertices of one BLAS-order triangle
  float p0[3]; float pad0;
  float p1[3]; float pad1;
  float p2[3]; float pad2;
};
static_assert(sizeof(GTri) == 48, "GTri must be 48 B");

struct __align__(16) GTree {  // root of a TLAS / BLAS
  float bmin[3]; uint32_t root;       // root ref (a leaf ref if the tree is a single leaf / single instance)
  float bmax[3]; uint32_t node_base;  // first GNode of this tree in the shared node array
};

struct __align__(16) GMesh {
  GTree tree;          // 32 B
  uint32_t tri_base;   // first GTri
  uint32_t tri_count;
  uint32_t alias_base;
  uint32_t pad;
};

// What "entering" an instance needs, in one 96-B record (no dependent
// instance -> mesh fetch): WorldToObject, then the primitive's BLAS root.
struct __align__(16) GInstTrav {
  float w2o[12];                      // rows 1..4 x columns 1..3 of WorldToObject
  float bmin[3]; uint32_t root;       // BLAS root bounds (MeshPrimitive.Bounds) + root ref | sphere: unused
  float bmax[3]; uint32_t node_base;  // mesh's first GNode
  uint32_t tri_base;                  // mesh's first GTri
  uint32_t is_sphere;
  float radius;
  uint32_t identity;                  // mesh instance whose ObjectToWorld AND WorldToObject are exactly the identity
};
static_assert(sizeof(GInstTrav) == 96, "GInstTrav must be 96 B");

struct __align__(16) GInstHead {  // shading-side instance record (48 B)
  float bmin[3]; uint32_t kind_prim;  // bit31: sphere, low bits: mesh / sphere index
  float bmax[3]; int32_t material;    // -1 = none
  int32_t light;                      // -1 = none
  float light_pdf_area;               // MeshInstance.EvalPDF with tag=0 (Mesh.fs:300-304), host-precomputed; spheres: radius
  uint32_t pad[2];
};
static_assert(sizeof(GInstHead) == 48, "GInstHead must be 48 B");

// Small-TLAS fast path: for each of the 8 direction octants, the instances in the
// order the reference's TLAS walk reaches them (the order depends only on the signs
// of the ray direction), each with its world AABB: 32 B per entry.
struct __align__(16) GFlatInst {
  float bmin[3]; uint32_t slot;
  float bmax[3]; uint32_t direct_root;  // identity mesh instance: its BLAS root ref (the scan enters it without phase E); else 0xFFFFFFFF
};
constexpr uint32_t kFlatTlasMax = 16;  // use the ordered scan when the scene has at most this many instances

struct __align__(16) GMat43 { float m[12]; };  // rows 1..4 x columns 1..3 of a Matrix4x4

struct __align__(16) GMaterial { uint32_t type; float r, g, b; float p0, p1, pad0, pad1; };
struct __align__(16) GLight { float r, g, b; uint32_t two_sided; };
struct GAlias { int32_t alias; float prob; float pdf; };

struct GCamera {
  uint32_t type;
  float viewport_h;  // 2*tan(fovY*pi/360), evaluated on the host with libm exactly as Pinhole.fs:15
  float aspect, aperture, focus, push_forward;
  float c2w[12];
};

struct DScene {
  const GNode* nodes;        // TLAS nodes first, then every mesh's nodes
  GTree tlas;
  const GInstTrav* inst_trav;
  const GInstHead* inst_head;
  const GMat43* inst_w2o;
  const GMat43* inst_o2w;
  const GMesh* meshes;
  const GTri* tris;
  const GAlias* alias;
  const float* sphere_radii;
  const GMaterial* materials;
  const GLight* lights;
  const uint32_t* light_inst;
  uint32_t n_inst, n_light_inst;
  uint32_t all_finite;       // every box / vertex / matrix is finite: the fast slab path is exact (vecmath.cuh)
  const GFlatInst* flat_tlas;  // [8][n_inst] or nullptr when n_inst > kFlatTlasMax
  GCamera cam;
};

}  // namespace bn
