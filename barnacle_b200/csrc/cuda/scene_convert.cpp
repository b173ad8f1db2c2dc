// Host-side conversion of the reference-layout scene into the traversal layout.
// Compiled by g++ with -ffp-contract=off: the one arithmetic result produced
// here (MeshInstance.EvalPDF for tag 0, Mesh.fs:300-304) must have the bits the
// reference's op sequence gives.
#include "scene_convert.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "../host/host_math.hpp"
#include "traverse_limits.h"

namespace bnconv {
namespace {

using bnhost::V3;

struct TreeOut {
  bn::GTree tree;
  int depth = 0;     // GNode levels incl. the pseudo chains of TLAS leaves
  int max_leaf = 0;
  bool finite = true;
};

bool finite3(const float* v) { return std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]); }

// Converts one reference BVH (preorder BnBVHNode array) into GNodes appended at
// out[node_base...].  BLAS (instances == nullptr): a leaf becomes a leaf ref
// (count, first triangle).  TLAS: every leaf ref is ONE instance whose world AABB
// sits in its parent, so the traversal tests it like any child box; a reference
// leaf holding k > 1 instances becomes a chain of k-1 pseudo nodes
//     P1{ left = inst f, right = P2 } , P2{ left = inst f+1, right = ... }
// with axis = 3 ("always left first") and +-FLT_MAX bounds for the chain links,
// which visits the instances in slot order, each against the then-current t —
// exactly the loop at Aggregate/BVH.fs:49-50 (a link fails only when t < 1e-3, and
// then every instance test would fail too because tMin >= 1e-3).
bool build_tree(const BnBVHNode* n, uint32_t count, uint32_t item_count, uint32_t node_base, uint32_t item_base, const BnInstance* instances,
                std::vector<bn::GNode>& out, TreeOut& t, std::string& err, const char* what) {
  if (count == 0) { err = std::string(what) + ": empty BVH"; return false; }
  std::vector<int> gidx(count, -1);
  int next = 0;
  for (uint32_t i = 0; i < count; ++i) {
    if (!n[i].is_leaf) gidx[i] = next++;
    if (!finite3(n[i].bounds_min) || !finite3(n[i].bounds_max)) t.finite = false;
  }
  out.resize(node_base + (size_t)next);
  // returns the ref for reference leaf i and the extra GNode levels it adds
  auto leaf_ref = [&](uint32_t i, uint32_t& ref, int& extra) {
    int c = n[i].count, first = n[i].right_or_offset;
    extra = 0;
    if (c <= 0) { err = std::string(what) + ": leaf item count out of range"; return false; }
    if (first < 0 || (uint32_t)first + (uint32_t)c > item_count) { err = std::string(what) + ": leaf item range out of bounds"; return false; }
    if (c > t.max_leaf) t.max_leaf = c;
    if (!instances) {
      // refs are ABSOLUTE (scene-wide triangle index): 3-bit count, 27-bit first
      if (c > bn::kMaxLeafCount || (uint64_t)item_base + (uint64_t)first >= bn::kMaxLeafFirst) { err = std::string(what) + ": leaf too large for the 3-bit count / 27-bit offset encoding"; return false; }
      ref = bn::kLeafBit | ((uint32_t)c << 27) | (item_base + (uint32_t)first);
      return true;
    }
    if ((uint32_t)first + (uint32_t)c > (1u << 30)) { err = "too many instances"; return false; }
    if (c == 1) { ref = bn::kLeafBit | (uint32_t)first; return true; }
    // chain of c-1 pseudo nodes, appended after the real interior nodes
    const size_t base = out.size();
    out.resize(base + (size_t)(c - 1));
    for (int k = 0; k < c - 1; ++k) {
      bn::GNode g;
      const BnInstance& a = instances[first + k];
      std::memcpy(g.lmin, a.bounds_min, 12); std::memcpy(g.lmax, a.bounds_max, 12);
      g.left = bn::kLeafBit | (uint32_t)(first + k);
      if (k == c - 2) {
        const BnInstance& b = instances[first + k + 1];
        std::memcpy(g.rmin, b.bounds_min, 12); std::memcpy(g.rmax, b.bounds_max, 12);
        g.right = bn::kLeafBit | (uint32_t)(first + k + 1);
      } else {
        for (int d = 0; d < 3; ++d) { g.rmin[d] = -3.402823466e38f; g.rmax[d] = 3.402823466e38f; }
        g.right = (uint32_t)(base + (size_t)k + 1);
      }
      g.axis = 3;
      g.pad = 0;
      out[base + (size_t)k] = g;
    }
    ref = (uint32_t)base;
    extra = c - 1;
    return true;
  };
  auto child_ref = [&](uint32_t c, uint32_t& ref, int& extra) {
    if (n[c].is_leaf) return leaf_ref(c, ref, extra);
    ref = node_base + (uint32_t)gidx[c];  // ABSOLUTE GNode index
    extra = 0;
    return true;
  };
  // explicit stack (preorder array: left = i+1, right = RightChild)
  std::vector<std::pair<uint32_t, int>> stack{{0u, 1}};
  std::vector<char> seen(count, 0);
  while (!stack.empty()) {
    auto [i, depth] = stack.back();
    stack.pop_back();
    if (i >= count || seen[i]) { err = std::string(what) + ": malformed BVH (node visited twice or out of range)"; return false; }
    seen[i] = 1;
    if (depth > t.depth) t.depth = depth;
    if (n[i].is_leaf) continue;
    int r = n[i].right_or_offset;
    uint32_t l = i + 1;
    if (l >= count || r <= (int)l || (uint32_t)r >= count) { err = std::string(what) + ": malformed BVH (child index)"; return false; }
    if (n[i].split_axis < 0 || n[i].split_axis > 2) { err = std::string(what) + ": split axis out of range"; return false; }
    bn::GNode g;
    std::memcpy(g.lmin, n[l].bounds_min, 12); std::memcpy(g.lmax, n[l].bounds_max, 12);
    std::memcpy(g.rmin, n[r].bounds_min, 12); std::memcpy(g.rmax, n[r].bounds_max, 12);
    int el = 0, er = 0;
    if (!child_ref(l, g.left, el) || !child_ref((uint32_t)r, g.right, er)) return false;
    if (depth + 1 + std::max(el, er) > t.depth) t.depth = depth + 1 + std::max(el, er);
    g.axis = (uint32_t)n[i].split_axis;
    g.pad = 0;
    out[node_base + (size_t)gidx[i]] = g;
    stack.push_back({l, depth + 1});
    stack.push_back({(uint32_t)r, depth + 1});
  }
  std::memcpy(t.tree.bmin, n[0].bounds_min, 12);
  std::memcpy(t.tree.bmax, n[0].bounds_max, 12);
  t.tree.node_base = node_base;
  if (n[0].is_leaf) {
    int extra = 0;
    if (!leaf_ref(0, t.tree.root, extra)) return false;
    t.depth = 1 + extra;
  } else {
    t.tree.root = node_base;
  }
  if (out.size() >= (1u << 30)) { err = "too many BVH nodes"; return false; }
  return true;
}

// ---- 4-wide nodes of the fast path (device_scene.h: GWide) -----------------------------------------------------------
// Built from the binary GNode array above, so TLAS (with its pseudo chains) and BLAS go through the same code.
struct WideBuilder {
  const std::vector<bn::GNode>& nodes;
  std::vector<bn::GWide>& wide;
  bool ok = true;   // false: some box is not finite / not ordered / not inside its parent's — the binary path is used
  int depth = 0;    // 4-wide levels of the tree being collapsed

  static constexpr float kBig = 3.402823466e38f;
  static bool is_link(const float* lo, const float* hi) { return lo[0] == -kBig && hi[0] == kBig; }  // chain link of a pseudo node
  bool box_ok(const float* lo, const float* hi) {
    if (is_link(lo, hi)) return true;
    for (int a = 0; a < 3; ++a)
      if (!std::isfinite(lo[a]) || !std::isfinite(hi[a]) || !(lo[a] <= hi[a])) return false;
    return true;
  }
  static bool inside(const float* olo, const float* ohi, const float* ilo, const float* ihi) {
    for (int a = 0; a < 3; ++a)
      if (!(olo[a] <= ilo[a] && ihi[a] <= ohi[a])) return false;
    return true;
  }
  void set_slot(bn::GWide& w, int k, const float* lo, const float* hi, uint32_t ref) {
    if (!box_ok(lo, hi)) ok = false;
    for (int a = 0; a < 3; ++a) { w.lo[a][k] = lo[a]; w.hi[a][k] = hi[a]; }
    w.ref[k] = ref;
  }
  // instances of the pseudo chain that starts at interior node c (axis 3): 2 .. k
  int chain_length(uint32_t c) const {
    int n = 1;
    while (!(c & bn::kLeafBit) && nodes[c].axis == 3u) { ++n; c = nodes[c].right; }
    return n;
  }
  uint32_t child(uint32_t ref, int level) { return (ref & bn::kLeafBit) ? ref : collapse(ref, level); }

  // `ref`: an interior GNode ref.  Returns the index of the GWide that stands for it.
  uint32_t collapse(uint32_t ref, int level) {
    const uint32_t idx = (uint32_t)wide.size();
    wide.emplace_back();
    if (level + 1 > depth) depth = level + 1;
    bn::GWide w;
    for (int a = 0; a < 3; ++a)
      for (int k = 0; k < 4; ++k) { w.lo[a][k] = kBig; w.hi[a][k] = -kBig; }
    for (int k = 0; k < 4; ++k) w.ref[k] = bn::kWideEmpty;
    w.pad[0] = w.pad[1] = 0;
    const bn::GNode P = nodes[ref];
    if (P.axis == 3u) {
      // TLAS leaf holding several instances (a chain of pseudo nodes): up to four of them in ONE wide node, in slot
      // order (every axis 3); a longer leaf continues in slot 3 behind the chain's always-passing link box
      uint32_t c = ref;
      int k = 0;
      for (;;) {
        const bn::GNode C = nodes[c];
        if (C.axis != 3u || !(C.left & bn::kLeafBit)) { ok = false; break; }
        set_slot(w, k++, C.lmin, C.lmax, C.left);
        if (C.right & bn::kLeafBit) { set_slot(w, k++, C.rmin, C.rmax, C.right); break; }
        if (k == 3) { set_slot(w, 3, C.rmin, C.rmax, collapse(C.right, level + 1)); break; }
        c = C.right;
      }
      w.axes = 3u | (3u << 2) | (3u << 4);
    } else {
      uint32_t axes = P.axis;
      for (int g = 0; g < 2; ++g) {
        const uint32_t c = g == 0 ? P.left : P.right;
        const float* clo = g == 0 ? P.lmin : P.rmin;
        const float* chi = g == 0 ? P.lmax : P.rmax;
        uint32_t axis = 3u;
        if ((c & bn::kLeafBit) || (nodes[c].axis == 3u && chain_length(c) > 2)) {
          // a leaf, or a TLAS leaf of 3+ instances kept behind its own box (one more node visit only when that box passes)
          for (uint32_t q = c; !(q & bn::kLeafBit); q = nodes[q].right) {  // ... which must hold every instance of the chain
            if (!inside(clo, chi, nodes[q].lmin, nodes[q].lmax)) ok = false;
            if ((nodes[q].right & bn::kLeafBit) && !inside(clo, chi, nodes[q].rmin, nodes[q].rmax)) ok = false;
          }
          set_slot(w, 2 * g, clo, chi, child(c, level + 1));
        } else {
          const bn::GNode C = nodes[c];
          if (!inside(clo, chi, C.lmin, C.lmax) || !inside(clo, chi, C.rmin, C.rmax)) ok = false;
          set_slot(w, 2 * g, C.lmin, C.lmax, child(C.left, level + 1));
          set_slot(w, 2 * g + 1, C.rmin, C.rmax, child(C.right, level + 1));
          axis = C.axis;
        }
        axes |= axis << (2 + 2 * g);
      }
      w.axes = axes;
    }
    // visiting order per direction octant: "left first iff dir[axis] > 0" (BVH.fs:51-56, Mesh.fs:235-240); axis 3 = slot order
    w.flips = 0;
    for (uint32_t oct = 0; oct < 8; ++oct) {
      const uint32_t sg = oct | 8u;
      const uint32_t fl = (((sg >> ((w.axes >> 2) & 3u)) & 1u) ? 0u : 1u) | (((sg >> ((w.axes >> 4) & 3u)) & 1u) ? 0u : 2u) | (((sg >> (w.axes & 3u)) & 1u) ? 0u : 4u);
      w.flips |= fl << (3u * oct);
    }
    wide[idx] = w;
    return idx;
  }
};

void mat43(const float* m, bn::GMat43& o) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 3; ++c) o.m[r * 3 + c] = m[r * 4 + c];
}

}  // namespace

bool convert_scene(const BnSceneDesc& d, ConvertedScene& out, std::string& err) {
  if (!d.tlas_nodes || !d.instances || d.instance_count == 0) { err = "scene has no instances / TLAS"; return false; }
  if ((d.mesh_count && !d.meshes) || (d.vertex_count && !d.vertices) || (d.triangle_count && !d.triangles) || (d.blas_node_count && !d.blas_nodes) ||
      (d.alias_count && !d.alias) || (d.sphere_count && !d.sphere_radii) || (d.material_count && !d.materials) || (d.light_count && !d.lights) ||
      (d.light_instance_count && !d.light_instances)) {
    err = "scene description has a NULL array with a non-zero count";
    return false;
  }
  TreeOut tl;
  if (!build_tree(d.tlas_nodes, d.tlas_node_count, d.instance_count, 0, 0, d.instances, out.nodes, tl, err, "TLAS")) return false;
  out.tlas = tl.tree;
  out.all_finite = tl.finite;
  int max_blas_depth = 0;
  // meshes
  out.meshes.resize(d.mesh_count);
  out.tris.resize(d.triangle_count);
  for (uint32_t m = 0; m < d.mesh_count; ++m) {
    const BnMesh& mm = d.meshes[m];
    if ((uint64_t)mm.vertex_offset + mm.vertex_count > d.vertex_count || (uint64_t)mm.tri_offset + mm.tri_count > d.triangle_count ||
        (uint64_t)mm.node_offset + mm.node_count > d.blas_node_count || (uint64_t)mm.alias_offset + mm.tri_count > d.alias_count || mm.tri_count == 0) {
      err = "mesh slice out of range";
      return false;
    }
    // AliasTable.Sample follows `alias` into the SAME mesh's table and triangle slice (light_sampler_sample, shade.cuh)
    for (uint32_t t = 0; t < mm.tri_count; ++t) {
      const int32_t a = d.alias[(size_t)mm.alias_offset + t].alias;
      if (a < 0 || (uint32_t)a >= mm.tri_count) { err = "alias entry points outside its mesh's triangle slice"; return false; }
    }
    TreeOut bt;
    if (!build_tree(d.blas_nodes + mm.node_offset, mm.node_count, mm.tri_count, (uint32_t)out.nodes.size(), mm.tri_offset, nullptr, out.nodes, bt, err, "BLAS")) return false;
    if (bt.depth > max_blas_depth) max_blas_depth = bt.depth;
    if (!bt.finite) out.all_finite = false;
    bn::GMesh& g = out.meshes[m];
    g.tree = bt.tree;
    g.tri_base = mm.tri_offset;
    g.tri_count = mm.tri_count;
    g.alias_base = mm.alias_offset;
    g.pad = 0;
    const float* v = d.vertices + (size_t)mm.vertex_offset * 3;
    for (uint32_t t = 0; t < mm.tri_count; ++t) {
      const int32_t* ix = d.triangles + ((size_t)mm.tri_offset + t) * 3;
      bn::GTri& gt = out.tris[mm.tri_offset + t];
      std::memset(&gt, 0, sizeof gt);
      for (int k = 0; k < 3; ++k) {
        if (ix[k] < 0 || (uint32_t)ix[k] >= mm.vertex_count) { err = "triangle vertex index out of range"; return false; }
        float* dst = k == 0 ? gt.p0 : (k == 1 ? gt.p1 : gt.p2);
        std::memcpy(dst, v + (size_t)ix[k] * 3, 12);
      }
    }
  }
  for (size_t k = 0; k < (size_t)d.vertex_count * 3; ++k)
    if (!std::isfinite(d.vertices[k])) { out.all_finite = false; break; }
  out.max_stack = tl.depth + max_blas_depth + 2;
  if (out.max_stack > bn::kStackSize) {
    err = "scene needs a traversal stack of " + std::to_string(out.max_stack) + " entries (limit " + std::to_string(bn::kStackSize) + ")";
    return false;
  }
  // instances
  out.inst_head.resize(d.instance_count);
  out.inst_trav.resize(d.instance_count);
  out.inst_w2o.resize(d.instance_count);
  out.inst_o2w.resize(d.instance_count);
  for (uint32_t i = 0; i < d.instance_count; ++i) {
    const BnInstance& in = d.instances[i];
    bn::GInstHead& h = out.inst_head[i];
    std::memset(&h, 0, sizeof h);
    std::memcpy(h.bmin, in.bounds_min, 12);
    std::memcpy(h.bmax, in.bounds_max, 12);
    if (in.material_id >= (int)d.material_count || in.light_id >= (int)d.light_count) { err = "instance material/light index out of range"; return false; }
    h.material = in.material_id < 0 ? -1 : in.material_id;
    h.light = in.light_id < 0 ? -1 : in.light_id;
    mat43(in.world_to_object, out.inst_w2o[i]);
    mat43(in.object_to_world, out.inst_o2w[i]);
    bn::GInstTrav& tv = out.inst_trav[i];
    std::memset(&tv, 0, sizeof tv);
    std::memcpy(tv.w2o, out.inst_w2o[i].m, sizeof tv.w2o);
    if (!finite3(in.bounds_min) || !finite3(in.bounds_max)) out.all_finite = false;
    for (float v : tv.w2o) if (!std::isfinite(v)) out.all_finite = false;
    if (in.prim_kind == BN_PRIM_SPHERE) {
      if (in.prim_id >= d.sphere_count) { err = "instance sphere index out of range"; return false; }
      h.kind_prim = 0x80000000u | in.prim_id;
      h.light_pdf_area = d.sphere_radii[in.prim_id];
      tv.is_sphere = 1;
      tv.radius = d.sphere_radii[in.prim_id];
    } else if (in.prim_kind == BN_PRIM_MESH) {
      if (in.prim_id >= d.mesh_count) { err = "instance mesh index out of range"; return false; }
      h.kind_prim = in.prim_id;
      const bn::GMesh& gm = out.meshes[in.prim_id];
      std::memcpy(tv.bmin, gm.tree.bmin, 12); std::memcpy(tv.bmax, gm.tree.bmax, 12);
      tv.root = gm.tree.root; tv.wroot = gm.tree.root; tv.tri_base = gm.tri_base;
      // Identity instances (e.g. the Cornell-box walls): Vector3.Transform by I returns its argument
      // bit for bit for the rays the fast path takes (no zero direction component; the sign of a
      // zero in the origin cannot reach a result), and the BLAS root box IS the instance's world
      // box, already tested.  The traversal then skips the transform, the reciprocal and that test.
      static const float kIdentity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
      const bool same_box = std::memcmp(in.bounds_min, gm.tree.bmin, 12) == 0 && std::memcmp(in.bounds_max, gm.tree.bmax, 12) == 0;
      bool ident = same_box;
      for (int k = 0; k < 16; ++k)  // numeric compare: a -0.0 entry (Matrix4x4.Invert produces them) acts like +0.0 here
        ident = ident && in.object_to_world[k] == kIdentity[k] && in.world_to_object[k] == kIdentity[k];
      tv.identity = ident ? 1u : 0u;
      // MeshInstance.EvalPDF (Mesh.fs:300-304) for tag = 0:
      // Table[0].pdf / SurfaceArea(Transform(triangle 0, ObjectToWorld))
      const BnMesh& mm = d.meshes[in.prim_id];
      const bn::GTri& t0 = out.tris[mm.tri_offset];
      bnhost::M4 M;
      std::memcpy(M.m, in.object_to_world, 64);
      V3 p0 = bnhost::transform_point({t0.p0[0], t0.p0[1], t0.p0[2]}, M);
      V3 p1 = bnhost::transform_point({t0.p1[0], t0.p1[1], t0.p1[2]}, M);
      V3 p2 = bnhost::transform_point({t0.p2[0], t0.p2[1], t0.p2[2]}, M);
      float area = 0.5f * bnhost::length(bnhost::cross(p1 - p0, p2 - p0));
      h.light_pdf_area = d.alias[mm.alias_offset].pdf / area;
    } else {
      err = "unknown primitive kind";
      return false;
    }
  }
  // 4-wide nodes of the fast path: the TLAS first, then every mesh.  Any box that is not finite / ordered / contained
  // in its parent's switches the whole scene back to the binary nodes (a foreign host may hand over any tree).
  {
    WideBuilder wb{out.nodes, out.wide};
    auto collapse_root = [&](uint32_t root, int& depth) {
      wb.depth = 0;
      const uint32_t r = (root & bn::kLeafBit) ? root : wb.collapse(root, 0);
      depth = wb.depth;
      return r;
    };
    int dt = 0, db_max = 0;
    out.tlas_wroot = collapse_root(out.tlas.root, dt);
    std::vector<uint32_t> mesh_wroot(d.mesh_count);
    for (uint32_t m = 0; m < d.mesh_count; ++m) {
      int db = 0;
      mesh_wroot[m] = collapse_root(out.meshes[m].tree.root, db);
      if (db > db_max) db_max = db;
    }
    out.max_stack_wide = 3 * (dt + db_max) + 2;   // every wide node on the current path leaves at most 3 entries
    out.inst_wroot.assign(d.instance_count, bn::kWideEmpty);
    for (uint32_t i = 0; i < d.instance_count; ++i)
      if (d.instances[i].prim_kind == BN_PRIM_MESH) out.inst_wroot[i] = mesh_wroot[d.instances[i].prim_id];
    if (!wb.ok || !out.all_finite || out.max_stack_wide > bn::kStackSize || out.wide.size() >= (1u << 30)) {
      out.wide.clear();
      out.inst_wroot.clear();
      out.max_stack_wide = 0;
    }
    for (uint32_t i = 0; i < d.instance_count; ++i)
      out.inst_trav[i].wroot = out.wide.empty() ? out.inst_trav[i].root : out.inst_wroot[i];
    // Bank swizzle (device_scene.h: kWideSwizzle): node i keeps its logical 16-B chunk j at physical chunk j ^ (i & 7)
    for (size_t i = 0; BN_WIDE_SWIZZLE && i < out.wide.size(); ++i) {
      const bn::GWide lg = out.wide[i];
      const unsigned char* src = reinterpret_cast<const unsigned char*>(&lg);
      unsigned char* dst = reinterpret_cast<unsigned char*>(&out.wide[i]);
      for (unsigned j = 0; j < 8; ++j) std::memcpy(dst + 16u * (j ^ (unsigned)(i & 7u)), src + 16u * j, 16);
    }
  }
  // Small-TLAS ordered scan (traverse.cuh): instance order of the reference's walk for each octant.
  // The walk visits left first iff dir[splitAxis] > 0 (Aggregate/BVH.fs:51-56) and leaf items in
  // slot order (:49-50), so the order is a function of the three sign bits only.
  if (d.instance_count <= bn::kFlatTlasMax) {
    out.flat_tlas.resize((size_t)8 * d.instance_count);
    for (uint32_t oct = 0; oct < 8; ++oct) {
      std::vector<uint32_t> order;
      std::vector<uint32_t> stack{0u};
      while (!stack.empty()) {
        const uint32_t i = stack.back();
        stack.pop_back();
        const BnBVHNode& nd = d.tlas_nodes[i];
        if (nd.is_leaf) {
          for (int k = 0; k < nd.count; ++k) order.push_back((uint32_t)(nd.right_or_offset + k));
        } else if ((oct >> nd.split_axis) & 1u) {  // dir[axis] > 0: push right, then left (popped first)
          stack.push_back((uint32_t)nd.right_or_offset);
          stack.push_back(i + 1);
        } else {
          stack.push_back(i + 1);
          stack.push_back((uint32_t)nd.right_or_offset);
        }
      }
      if (order.size() != d.instance_count) { err = "TLAS does not cover every instance exactly once"; return false; }
      for (uint32_t k = 0; k < d.instance_count; ++k) {
        bn::GFlatInst& f = out.flat_tlas[(size_t)oct * d.instance_count + k];
        const BnInstance& in = d.instances[order[k]];
        // near / far plane per axis for THIS octant (bit a of oct = dir[a] > 0: near = min): the scan's slab test needs no
        // min / max per axis — same values as Min/MaxNative(t0, t1) for every ray the fast path takes (device_scene.h)
        for (int a = 0; a < 3; ++a) {
          const bool pos = ((oct >> a) & 1u) != 0u;
          f.bmin[a] = pos ? in.bounds_min[a] : in.bounds_max[a];
          f.bmax[a] = pos ? in.bounds_max[a] : in.bounds_min[a];
        }
        f.slot = order[k];
        f.direct_root = out.inst_trav[order[k]].identity ? out.inst_trav[order[k]].wroot : 0xFFFFFFFFu;  // (use_binary_nodes() re-points it)
      }
    }
  }
  out.light_inst.assign(d.light_instances, d.light_instances + d.light_instance_count);
  for (uint32_t li : out.light_inst)
    if (li >= d.instance_count || d.instances[li].light_id < 0) { err = "light instance list is inconsistent"; return false; }
  out.alias.resize(d.alias_count);
  for (uint32_t i = 0; i < d.alias_count; ++i) {
    if (d.alias[i].alias < 0) { err = "alias entry out of range"; return false; }
    out.alias[i] = {d.alias[i].alias, d.alias[i].prob, d.alias[i].pdf};
  }
  out.sphere_radii.assign(d.sphere_radii, d.sphere_radii + d.sphere_count);
  out.materials.resize(d.material_count);
  for (uint32_t i = 0; i < d.material_count; ++i) {
    const BnMaterial& m = d.materials[i];
    if (m.type > BN_MAT_PBR) { err = "unknown material type"; return false; }
    out.materials[i] = {m.type, m.base_color[0], m.base_color[1], m.base_color[2], m.p0, m.p1, 0.f, 0.f};
  }
  out.lights.resize(d.light_count);
  for (uint32_t i = 0; i < d.light_count; ++i) out.lights[i] = {d.lights[i].emission[0], d.lights[i].emission[1], d.lights[i].emission[2], d.lights[i].two_sided ? 1u : 0u};
  // camera
  if (d.camera.type > BN_CAM_THIN_LENS) { err = "unknown camera type"; return false; }
  convert_camera(d.camera, out.cam);
  return true;
}

void convert_camera(const BnCamera& c, bn::GCamera& cam) {
  cam.type = c.type;
  cam.viewport_h = 2.f * tanf(c.fov_y * 3.14159274101257324f / 360.f);  // Pinhole.fs:15
  cam.aspect = c.aspect_ratio;
  cam.aperture = c.aperture;
  cam.focus = c.focus_distance;
  cam.push_forward = c.push_forward;
  bn::GMat43 cm;
  mat43(c.camera_to_world, cm);
  std::memcpy(cam.c2w, cm.m, sizeof cm.m);
}

// A/B switch (BN_BINARY_NODES): drop the 4-wide nodes and point everything that names a BLAS root back at the binary tree.
void use_binary_nodes(ConvertedScene& cs) {
  cs.wide.clear();
  cs.inst_wroot.clear();
  cs.max_stack_wide = 0;
  for (bn::GInstTrav& t : cs.inst_trav) t.wroot = t.root;
  for (bn::GFlatInst& f : cs.flat_tlas)
    if (f.direct_root != 0xFFFFFFFFu) f.direct_root = cs.inst_trav[f.slot].root;
}

}  // namespace bnconv
