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

// 4-wide node of the fast traversal path: a binary node P collapsed with its two children L, R.  Slots 0, 1 hold L's
// children (or L itself in slot 0 when L is a leaf), slots 2, 3 R's (or R in slot 2); an empty slot has an inverted box
// (lo = +FLT_MAX, hi = -FLT_MAX: the slab test fails for every ray) and ref kWideEmpty.  The reference's stack walk
// reaches the four grandchildren in the order [near(P).near, near(P).far, far(P).near, far(P).far], i.e. group L before
// group R iff dir[axis_P] > 0, slot 0 before slot 1 iff dir[axis_L] > 0, slot 2 before slot 3 iff dir[axis_R] > 0; an
// axis of 3 means "in slot order" (leaf children, TLAS leaves holding several instances).  Skipping the boxes of L and R
// cannot change a fast-path result: a grandchild's box lies inside its parent's, and the slab test is monotone in the
// box for finite non-NaN operands, so "grandchild passes" implies "child passes" with the same t (scene_convert.cpp
// checks the containment and falls back to the binary nodes otherwise).
// Planes are stored per axis for the four slots (SoA), so that the near / far plane of an axis is ONE 16-B load whose
// address depends on the sign of the ray direction — no min / max / select per box: the hi planes sit 64 B after the lo
// planes, so "near" / "far" is bit 6 of the address (the array is 128-B aligned).
// BN_WIDE_SWIZZLE (compile-time, default off): in memory the eight 16-B chunks of node i are XOR-swizzled — logical chunk j
// sits at physical chunk j ^ (i & 7).  All lanes of a warp read the SAME logical chunk of DIFFERENT nodes at once, and the
// L1 data array serves one lane per cycle when every lane's 16 bytes sit at the same offset of their 128-B lines
// (profiles/r02_l1_gather_microbench.txt: 8 x LDG.128 of random 128-B records run at one lane-load per cycle per SM
// whatever the number of active lanes); the swizzle spreads a warp's loads over the eight bank groups.  Measured on the
// B200 (profiles/r02_ab_session4_swizzle.log): the L1 data pipe of the extend launches falls from 80-83 % to 74-75 % busy —
// and the frame time does not move (C1 -1.0 %, C2 -1.4 %, C3 +0.8 %, C4 +1.8 %): the kernel is bound by instruction issue
// (76 %), not by L1, and the swizzle costs two instructions per node step.  Once the paths were ordered between bounces
// (ray_sort.cuh) the kernel was no longer bound by issue slots and the relieved L1 pipe showed (profiles/r02_ab_session26_*.log:
// C2 extend -2.7 %, C4 frame -2.3 %, C1 / C3 equal): ON by default since.  The struct below is the LOGICAL layout.
#ifndef BN_WIDE_SWIZZLE
#define BN_WIDE_SWIZZLE 1
#endif
struct __align__(128) GWide {
  float lo[3][4];    //   +0: lo.x of slots 0..3 | +16: lo.y | +32: lo.z
  uint32_t ref[4];   //  +48: child refs (same encoding as GNode.left / right; interior = ABSOLUTE GWide index)
  float hi[3][4];    //  +64: hi.x               | +80: hi.y | +96: hi.z
  uint32_t flips;    // +112: for each direction octant o = (dx>0) | (dy>0)<<1 | (dz>0)<<2, three bits at 3*o:
                     //       bit 0 = slot 1 before slot 0, bit 1 = slot 3 before slot 2, bit 2 = group R before group L
  uint32_t axes;     // +116: axis_P | axis_L << 2 | axis_R << 4 (what `flips` was derived from; 3 = "in slot order")
  uint32_t pad[2];
};
static_assert(sizeof(GWide) == 128, "GWide must be 128 B");
constexpr uint32_t kWideEmpty = 0xFFFFFFFFu;

// BN_TRI64 (switch, off): the record padded to 64 B so that the closest-hit kernel fetches a triangle with TWO 256-bit loads
// instead of three 128-bit ones (a gathered load costs the L1 data pipe a lane-cycle whatever its width,
// profiles/r02_l1_gather_microbench.txt).  Measured on the ordered kernel (profiles/r02_ab_session28_*.log): C2 extend
// 24.87 -> 25.03 ms, C1 / C3 / C4 within 0.2 % — the wavefronts saved are lost again to a third more triangle bytes in L1.
#ifndef BN_TRI64
#define BN_TRI64 0
#endif
#if BN_TRI64
struct __align__(32) GTri {  // pre-gathered vertices of one BLAS-order triangle
  float p0[3]; float pad0;
  float p1[3]; float pad1;
  float p2[3]; float pad2;
  float pad3[4];
};
static_assert(sizeof(GTri) == 64, "GTri must be 64 B");
#else
struct __align__(16) GTri {  // pre-gathered vertices of one BLAS-order triangle
  float p0[3]; float pad0;
  float p1[3]; float pad1;
  float p2[3]; float pad2;
};
static_assert(sizeof(GTri) == 48, "GTri must be 48 B");
#endif

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
  float bmax[3]; uint32_t wroot;      // root ref in the 4-wide node array (fast path); == root when the BLAS is a single leaf
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
// The box is stored as the octant's NEAR planes (bmin[a] = world min if dir[a] > 0 in this octant, else world max) and
// FAR planes (bmax): with lo <= hi, finite operands and a finite non-zero 1/d of that sign, (near - o) * inv is exactly
// Min(t0, t1) and (far - o) * inv exactly Max(t0, t1) of the reference's slab test (monotone rounding), up to the sign
// of a zero, which the Max with 1e-3 / the comparison against tMin >= 1e-3 cannot see.
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

// Grid of the path re-ordering between bounces (ray_sort.cuh): cell = clamp(int((p - lo) * scale), 0, 2^m - 1) per axis
struct SortGrid { float lo[3]; float scale[3]; uint32_t cell_major; uint32_t pad; };  // cell_major: key = cell above octant (scenes behind a small TLAS)

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
  const GWide* wide;           // 4-wide nodes of the fast path (TLAS first, then every mesh), or nullptr: binary fast path
  uint32_t tlas_wroot;         // TLAS root ref in `wide`
  uint32_t refill_min;         // closest-hit refill threshold of the traversal loop for this scene (0: the compile-time default, traverse.cuh)
  GCamera cam;
  SortGrid sort_grid;
};

}  // namespace bn
