// Host-side scene builder: everything the managed host does BEFORE the timed
// region of Scene.Render (Extensions/Scene/Render.fs:11-14) — JSON load
// (Extensions/Scene/Loader.fs), Scene.Traverse (Base/Scene.fs:43-61),
// MeshPrimitive construction incl. BLAS + alias table (Mesh.fs:119-186),
// BVHAggregate's TLAS (Aggregate/BVH.fs:9), UniformLightSampler's light list
// (Base/LightSampler.fs:7) — flattened into the BnSceneDesc arrays.
// It is a stand-in for the F# host (no dotnet in this image); it is NOT on the
// timed path and contains no ray tracing.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/barnacle_b200.h"
#include "host_math.hpp"
#include "json.hpp"

using namespace bnhost;

namespace bnhost {
void set_error(const std::string& msg);  // error.cpp
}

// ---------------------------------------------------------------------------
// BVHNode.Build — binned SAH (Util/BVH.fs:99-247)
// ---------------------------------------------------------------------------
namespace {

constexpr int kSAHBinCount = 12;  // BVHBuildConfig.SAHBinCount  (Util/BVH.fs:101)
constexpr int kMaxLeafSize = 4;   // BVHBuildConfig.MaxLeafSize  (:104)
constexpr int kMaxDepth = 64;     // BVHBuildConfig.MaxDepth     (:107)

struct BuildItem {
  AABB box;
  uint32_t index;
};

struct Bin {
  AABB box = AABB::empty();
  int count = 0;
  float sah_cost() const { return (float)count * box.surface_area(); }  // Util/BVH.fs:123
};

BnBVHNode make_leaf(const AABB& b, int first, int count) {
  BnBVHNode n;
  std::memset(&n, 0, sizeof n);
  n.bounds_min[0] = b.lo.x; n.bounds_min[1] = b.lo.y; n.bounds_min[2] = b.lo.z;
  n.bounds_max[0] = b.hi.x; n.bounds_max[1] = b.hi.y; n.bounds_max[2] = b.hi.z;
  n.right_or_offset = first;
  n.is_leaf = 1;
  n.count = (int8_t)count;  // int8 field (Util/BVH.fs:69-70,79-81)
  return n;
}

// Emits nodes in preorder while recursing (equivalent to Tree.Build followed by
// Flatten, Util/BVH.fs:224-237: a node's index is taken on entry, the left
// subtree follows immediately, RightChild = index of the right subtree's root).
int build_rec(std::vector<BuildItem>& items, int first, int last, int depth, std::vector<BnBVHNode>& nodes) {
  const int count = last - first;
  std::vector<BuildItem> sub(items.begin() + first, items.begin() + last);  // :130 (a copy)

  AABB bounds = AABB::empty();
  for (int i = 0; i < count; ++i) bounds = unite(bounds, sub[i].box);

  const int index = (int)nodes.size();
  if (count <= kMaxLeafSize || depth >= kMaxDepth) {  // :138
    nodes.push_back(make_leaf(bounds, first, count));
    return index;
  }

  AABB cb = AABB::empty();
  for (int i = 0; i < count; ++i) cb = unite(cb, sub[i].box.centroid());
  const int axis = cb.split_axis();
  const float extent = cb.diagonal()[axis];

  int mid;
  if (extent == 0.f) {  // :150-156 — median split, no reordering
    mid = first + count / 2;
  } else {
    Bin bins[kSAHBinCount];
    auto bin_of = [&](const BuildItem& it) {  // :163-169
      float f = ((float)kSAHBinCount * (it.box.centroid()[axis] - cb.lo[axis])) / extent;
      int b = (int)f;
      return std::min(b, kSAHBinCount - 1);
    };
    for (int i = 0; i < count; ++i) {
      Bin& b = bins[bin_of(sub[i])];
      b.box = unite(b.box, sub[i].box);
      b.count += 1;
    }
    constexpr int nc = kSAHBinCount - 1;
    float costs[nc];
    for (float& c : costs) c = 0.f;
    Bin left, right;
    for (int i = 0; i < nc; ++i) {  // :187-193
      left.box = unite(left.box, bins[i].box);
      left.count += bins[i].count;
      costs[i] = costs[i] + left.sah_cost();
      right.box = unite(right.box, bins[nc - i].box);
      right.count += bins[nc - i].count;
      costs[nc - 1 - i] = costs[nc - 1 - i] + right.sah_cost();
    }
    float min_cost = INFINITY;
    int best = 0;
    const float base = (float)count * cb.surface_area();
    for (int i = 0; i < nc; ++i) {  // :198-203
      float cost = costs[i] + base;
      if (cost < min_cost) { min_cost = cost; best = i; }
    }
    int l = first, r = last - 1;
    for (int i = 0; i < count; ++i) {  // :208-217 — right side filled from the end
      if (bin_of(sub[i]) > best) items[r--] = sub[i];
      else items[l++] = sub[i];
    }
    mid = l;
  }

  nodes.emplace_back();  // placeholder (Flatten adds default then overwrites, :232-234)
  build_rec(items, first, mid, depth + 1, nodes);
  int right_index = build_rec(items, mid, last, depth + 1, nodes);
  BnBVHNode n;
  std::memset(&n, 0, sizeof n);
  n.bounds_min[0] = bounds.lo.x; n.bounds_min[1] = bounds.lo.y; n.bounds_min[2] = bounds.lo.z;
  n.bounds_max[0] = bounds.hi.x; n.bounds_max[1] = bounds.hi.y; n.bounds_max[2] = bounds.hi.z;
  n.right_or_offset = right_index;
  n.is_leaf = 0;
  n.split_axis = (int8_t)axis;
  nodes[index] = n;
  return index;
}

// Returns nodes; perm[i] = original index of the item now at slot i.
thread_local int g_build_device = -1;  // >= 0: BVHs are built by bn_bvh_build on that CUDA device (bn_host_scene_load_ex)

std::vector<BnBVHNode> bvh_build(const std::vector<AABB>& boxes, std::vector<uint32_t>& perm) {
  if (g_build_device >= 0 && !boxes.empty()) {
    static_assert(sizeof(AABB) == 24, "AABB is 6 packed floats");
    std::vector<BnBVHNode> out(2 * boxes.size());
    perm.resize(boxes.size());
    const int cnt = bn_bvh_build(g_build_device, reinterpret_cast<const float*>(boxes.data()), (uint32_t)boxes.size(), out.data(), (uint32_t)out.size(),
                                 perm.data(), nullptr);
    if (cnt < 0) throw std::runtime_error(std::string("device BVH build failed: ") + bn_last_error());
    out.resize((size_t)cnt);
    return out;
  }
  std::vector<BuildItem> items(boxes.size());
  for (size_t i = 0; i < boxes.size(); ++i) items[i] = {boxes[i], (uint32_t)i};
  std::vector<BnBVHNode> nodes;
  build_rec(items, 0, (int)items.size(), 0, nodes);
  perm.resize(items.size());
  for (size_t i = 0; i < items.size(); ++i) perm[i] = items[i].index;
  return nodes;
}

// ---------------------------------------------------------------------------
// AliasTable ctor (Util/AliasTable.fs:14-51).  Restated as written, including
// the missing xN normalisation.
// ---------------------------------------------------------------------------
std::vector<BnAliasEntry> alias_build(const std::vector<float>& w) {
  const int n = (int)w.size();
  std::vector<BnAliasEntry> t(n);
  float sum = 0.f;
  for (float x : w) sum = sum + x;  // Array.sum: sequential fp32 accumulation
  if (sum == 0.f) {
    for (int i = 0; i < n; ++i) t[i] = {i, 1.f, 1.f / (float)n};
    return t;
  }
  for (int i = 0; i < n; ++i) t[i] = {i, w[i] / sum, w[i] / sum};
  std::vector<int> under, over;
  for (int i = 0; i < n; ++i) (t[i].prob > 1.f ? over : under).push_back(i);
  while (!under.empty() && !over.empty()) {
    int o = over.back(); over.pop_back();
    int u = under.back(); under.pop_back();
    t[o].prob = t[o].prob + t[u].prob - 1.f;
    t[u].alias = o;
    (t[o].prob > 1.f ? over : under).push_back(o);
  }
  for (int o : over) { t[o].alias = o; t[o].prob = 1.f; }
  for (int u : under) { t[u].alias = u; t[u].prob = 1.f; }
  return t;
}

// ---------------------------------------------------------------------------
// Scene objects
// ---------------------------------------------------------------------------
struct MeshData {
  std::vector<V3> vertices;
  std::vector<int32_t> tri;  // BLAS order after build
  std::vector<BnBVHNode> nodes;
  std::vector<BnAliasEntry> alias;
  std::vector<uint32_t> perm;
};

AABB tri_bounds(const MeshData& m, int t) {  // Mesh.fs:162-168
  V3 p0 = m.vertices.at(m.tri[t * 3]), p1 = m.vertices.at(m.tri[t * 3 + 1]), p2 = m.vertices.at(m.tri[t * 3 + 2]);
  return {min_native(min_native(p0, p1), p2), max_native(max_native(p0, p1), p2)};
}

void finish_mesh(MeshData& m) {
  const int nt = (int)m.tri.size() / 3;
  std::vector<AABB> boxes(nt);
  for (int t = 0; t < nt; ++t) boxes[t] = tri_bounds(m, t);
  m.nodes = bvh_build(boxes, m.perm);
  std::vector<int32_t> sorted(m.tri.size());
  for (int t = 0; t < nt; ++t)
    for (int k = 0; k < 3; ++k) sorted[t * 3 + k] = m.tri[m.perm[t] * 3 + k];
  m.tri.swap(sorted);
  // AliasTable over the (already permuted) triangles, Mesh.fs:181-186
  std::vector<float> area(nt);
  for (int t = 0; t < nt; ++t) {
    V3 p0 = m.vertices[m.tri[t * 3]], p1 = m.vertices[m.tri[t * 3 + 1]], p2 = m.vertices[m.tri[t * 3 + 2]];
    area[t] = 0.5f * length(cross(p1 - p0, p2 - p0));  // Triangle.SurfaceArea, Mesh.fs:84-87
  }
  m.alias = alias_build(area);
}

MeshData make_quad() {  // Mesh.fs:126-134
  MeshData m;
  m.vertices = {{-1, 0, -1}, {1, 0, -1}, {1, 0, 1}, {-1, 0, 1}};
  m.tri = {0, 1, 2, 0, 2, 3};
  finish_mesh(m);
  return m;
}
MeshData make_cube() {  // Mesh.fs:136-155
  MeshData m;
  m.vertices = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
  m.tri = {0, 2, 1, 0, 3, 2, 4, 5, 6, 4, 6, 7, 0, 1, 5, 0, 5, 4, 3, 7, 6, 3, 6, 2, 0, 4, 7, 0, 7, 3, 1, 2, 6, 1, 6, 5};
  finish_mesh(m);
  return m;
}

std::string trim(const std::string& s) {
  size_t b = s.find_first_not_of(" \t\r\n\v\f"), e = s.find_last_not_of(" \t\r\n\v\f");
  return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
}
std::vector<std::string> split_spaces(const std::string& s) {  // Split(' ', RemoveEmptyEntries)
  std::vector<std::string> out;
  size_t p = 0;
  while (p < s.size()) {
    size_t q = s.find(' ', p);
    if (q == std::string::npos) q = s.size();
    if (q > p) out.push_back(s.substr(p, q - p));
    p = q + 1;
  }
  return out;
}

MeshData load_obj(const std::string& path) {  // MeshPrimitive.Load, Mesh.fs:246-279
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open mesh file: " + path);
  MeshData m;
  std::string raw;
  while (std::getline(in, raw)) {
    std::string line = trim(raw);
    if (line.rfind("v ", 0) == 0) {
      auto tok = split_spaces(line.substr(2));
      if (tok.size() != 3) throw std::runtime_error("OBJ: vertex needs 3 coordinates");
      m.vertices.push_back({std::strtof(tok[0].c_str(), nullptr), std::strtof(tok[1].c_str(), nullptr), std::strtof(tok[2].c_str(), nullptr)});
    } else if (line.rfind("f ", 0) == 0) {
      auto tok = split_spaces(line.substr(2));
      if (tok.size() < 3) throw std::runtime_error("OBJ: face needs >= 3 vertices");
      auto index = [&](const std::string& t) {
        size_t s = t.find('/');
        int i = std::atoi((s == std::string::npos ? t : t.substr(0, s)).c_str());
        return i < 0 ? (int)m.vertices.size() + i : i - 1;
      };
      int i0 = index(tok[0]), i1 = index(tok[1]);
      for (size_t k = 2; k < tok.size(); ++k) {  // fan triangulation
        int i2 = index(tok[k]);
        m.tri.push_back(i0); m.tri.push_back(i1); m.tri.push_back(i2);
        i1 = i2;
      }
    }
  }
  finish_mesh(m);
  return m;
}

struct Primitive {
  int kind;  // BN_PRIM_*
  int id;    // mesh index or sphere index
};

struct InstanceObj {  // PrimitiveInstance (Base/Primitive.fs:100-141)
  Primitive prim;
  int material = -1, light = -1;
  M4 o2w = M4::identity(), w2o = M4::identity();
  AABB bounds = AABB::empty();
};

struct NodeObj {
  std::vector<M4> key_matrix;
  std::vector<float> key_time;
  std::vector<int> instances, children;
  bool has_camera = false;
};

V3 vec3_of(const bnjson::Value& a) {
  if (a.kind != bnjson::Value::Array || a.size() < 3) throw std::runtime_error("JSON: 3-vector expected");
  return {a[0].as_f32(), a[1].as_f32(), a[2].as_f32()};
}

}  // namespace

struct BnHostScene {
  std::vector<MeshData> meshes;
  std::vector<float> sphere_radii;
  std::vector<BnMaterial> materials;
  std::vector<BnLight> lights;
  BnHostSceneInfo info{};
  // flattened
  std::vector<BnBVHNode> tlas;
  std::vector<BnInstance> instances;
  std::vector<uint32_t> instance_perm;
  std::vector<uint32_t> light_instances;
  std::vector<BnMesh> mesh_recs;
  std::vector<float> vertices;
  std::vector<int32_t> triangles;
  std::vector<BnBVHNode> blas;
  std::vector<BnAliasEntry> alias;
  BnSceneDesc desc{};
};

namespace {

// Transform.Eval(t) (Base/Scene.fs:26-36): the last keyframe at or before t; the first keyframe when
// t precedes all of them; KeyFrame.Interpolate between it and the next one otherwise — also when t
// sits exactly ON a keyframe that is not the last (ratio 0), which, with Transform.Compose as written
// (host_math.hpp), is not that keyframe's matrix.  SURVEY §8(f) N4.
M4 eval_transform(const NodeObj& n, float t) {
  if (n.key_matrix.empty()) return M4::identity();
  int prev = -1;
  for (int i = (int)n.key_time.size() - 1; i >= 0; --i)
    if (n.key_time[i] <= t) { prev = i; break; }
  if (prev < 0) return n.key_matrix[0];
  if (prev == (int)n.key_matrix.size() - 1) return n.key_matrix[prev];
  return interpolate_keyframes(n.key_time[prev], n.key_matrix[prev], n.key_time[prev + 1], n.key_matrix[prev + 1], t);
}

void traverse(const std::vector<NodeObj>& nodes, int idx, float t, const M4& parent, std::vector<InstanceObj>& objs,
              std::vector<int>& order, M4& camera_to_world, const std::vector<MeshData>& meshes,
              const std::vector<float>& radii, int depth) {
  if (depth > 4096) throw std::runtime_error("scene graph too deep (cycle?)");
  const NodeObj& n = nodes.at(idx);
  M4 o2w = mul(eval_transform(n, t), parent);  // Scene.fs:44: local * parent
  if (n.has_camera) camera_to_world = o2w;
  for (int i : n.instances) {  // PrimitiveInstance.UpdateTransform, Primitive.fs:131-138
    InstanceObj& o = objs.at(i);
    o.o2w = o2w;
    M4 inv = M4::identity();
    invert(o2w, inv);
    o.w2o = inv;
    AABB pb;
    if (o.prim.kind == BN_PRIM_MESH) {
      const BnBVHNode& r = meshes[o.prim.id].nodes.at(0);  // MeshPrimitive.Bounds = BVHNodes[0].Bounds
      pb = {{r.bounds_min[0], r.bounds_min[1], r.bounds_min[2]}, {r.bounds_max[0], r.bounds_max[1], r.bounds_max[2]}};
    } else {
      float rad = radii[o.prim.id];  // Sphere.fs:79-80
      pb = {{-rad, -rad, -rad}, {rad, rad, rad}};
    }
    o.bounds = transform_aabb(pb, o2w);
    order.push_back(i);
  }
  for (int c : n.children) traverse(nodes, c, t, o2w, objs, order, camera_to_world, meshes, radii, depth + 1);
}

BnHostScene* build_scene(const std::string& text, const std::string& base_dir, float time) {
  auto root_v = bnjson::parse(text);
  const bnjson::Value& root = *root_v;
  auto need = [&](const char* k) -> const bnjson::Value& {
    const bnjson::Value* v = root.get(k);
    if (!v) throw std::runtime_error(std::string("scene JSON: missing '") + k + "'");
    return *v;
  };
  auto scene = std::make_unique<BnHostScene>();

  // ---- transforms (Loader.fs:14-48) : S * R * T, R = Rx * Ry * Rz
  struct TransformObj { std::vector<std::pair<float, M4>> keys; };
  std::vector<TransformObj> transforms;
  if (const bnjson::Value* ts = root.get("transforms")) {
    for (size_t i = 0; i < ts->size(); ++i) {
      TransformObj t;
      const bnjson::Value* kfs = (*ts)[i].get("keyframes");
      if (!kfs || kfs->size() == 0) throw std::runtime_error("KeyFrames cannot be empty");
      for (size_t k = 0; k < kfs->size(); ++k) {
        const bnjson::Value& kf = (*kfs)[k];
        M4 S = M4::identity(), R = M4::identity(), T = M4::identity();
        if (auto* s = kf.get("scale")) { V3 v = vec3_of(*s); S = create_scale(v.x, v.y, v.z); }
        if (auto* r = kf.get("rotation")) {
          V3 v = vec3_of(*r);
          R = mul(mul(create_rotation_x(v.x), create_rotation_y(v.y)), create_rotation_z(v.z));
        }
        if (auto* tr = kf.get("translation")) { V3 v = vec3_of(*tr); T = create_translation(v.x, v.y, v.z); }
        float tm = kf.get("time") ? kf.get("time")->as_f32() : 0.f;
        t.keys.emplace_back(tm, mul(mul(S, R), T));
      }
      std::stable_sort(t.keys.begin(), t.keys.end(), [](auto& a, auto& b) { return a.first < b.first; });
      transforms.push_back(std::move(t));
    }
  }

  // ---- primitives (Loader.fs:92-116)
  std::vector<Primitive> prims;
  {
    const bnjson::Value& ps = need("primitives");
    for (size_t i = 0; i < ps.size(); ++i) {
      const bnjson::Value& p = ps[i];
      const std::string& type = p.get("type") ? p.get("type")->as_string() : std::string();
      if (type == "sphere") {
        scene->sphere_radii.push_back(p.get("radius") ? p.get("radius")->as_f32() : 1.f);
        prims.push_back({BN_PRIM_SPHERE, (int)scene->sphere_radii.size() - 1});
      } else if (type == "quad") {
        scene->meshes.push_back(make_quad());
        prims.push_back({BN_PRIM_MESH, (int)scene->meshes.size() - 1});
      } else if (type == "cube") {
        scene->meshes.push_back(make_cube());
        prims.push_back({BN_PRIM_MESH, (int)scene->meshes.size() - 1});
      } else if (type == "mesh") {
        if (auto* uri = p.get("uri")) {
          std::string path = uri->as_string();
          if (!base_dir.empty() && !path.empty() && path[0] != '/') path = base_dir + "/" + path;
          scene->meshes.push_back(load_obj(path));
        } else if (p.get("vertices") && p.get("indices")) {
          MeshData m;
          const bnjson::Value& vs = *p.get("vertices");
          for (size_t k = 0; k + 2 < vs.size(); k += 3) m.vertices.push_back({vs[k].as_f32(), vs[k + 1].as_f32(), vs[k + 2].as_f32()});
          const bnjson::Value& is = *p.get("indices");
          for (size_t k = 0; k + 2 < is.size(); k += 3) { m.tri.push_back(is[k].as_int()); m.tri.push_back(is[k + 1].as_int()); m.tri.push_back(is[k + 2].as_int()); }
          for (int32_t ix : m.tri)
            if (ix < 0 || ix >= (int)m.vertices.size()) throw std::runtime_error("mesh index out of range");
          finish_mesh(m);
          scene->meshes.push_back(std::move(m));
        } else {
          throw std::runtime_error("Invalid mesh primitive");
        }
        prims.push_back({BN_PRIM_MESH, (int)scene->meshes.size() - 1});
      } else {
        throw std::runtime_error("Unknown primitive type: " + type);
      }
    }
  }

  // ---- materials (Loader.fs:50-73)
  if (const bnjson::Value* ms = root.get("materials")) {
    for (size_t i = 0; i < ms->size(); ++i) {
      const bnjson::Value& m = (*ms)[i];
      BnMaterial out{};
      V3 c = m.get("albedo") ? vec3_of(*m.get("albedo")) : V3{1, 1, 1};
      out.base_color[0] = c.x; out.base_color[1] = c.y; out.base_color[2] = c.z;
      float ior = m.get("ior") ? m.get("ior")->as_f32() : 1.5f;
      float rough = m.get("roughness") ? m.get("roughness")->as_f32() : 1.f;
      float metal = m.get("metallic") ? m.get("metallic")->as_f32() : 0.f;
      const std::string& type = m.get("type") ? m.get("type")->as_string() : std::string();
      if (type == "lambertian") out.type = BN_MAT_LAMBERTIAN;
      else if (type == "mirror") out.type = BN_MAT_MIRROR;
      else if (type == "dielectric") { out.type = BN_MAT_DIELECTRIC; out.p0 = ior; }
      else if (type == "pbr") {
        out.type = BN_MAT_PBR;
        out.p0 = metal < 0.f ? 0.f : (metal > 1.f ? 1.f : metal);  // Single.Clamp, PBR.fs:12
        float a = rough * rough;
        out.p1 = a > 1e-3f ? a : 1e-3f;                            // MathF.Max(r*r, 1e-3f), PBR.fs:11
      } else throw std::runtime_error("Unknown material type: " + type);
      scene->materials.push_back(out);
    }
  }
  // ---- lights (Loader.fs:75-90)
  if (const bnjson::Value* ls = root.get("lights")) {
    for (size_t i = 0; i < ls->size(); ++i) {
      const bnjson::Value& l = (*ls)[i];
      const std::string& type = l.get("type") ? l.get("type")->as_string() : std::string();
      if (type != "diffuse") throw std::runtime_error("Unknown light type: " + type);
      BnLight out{};
      V3 e = l.get("emission") ? vec3_of(*l.get("emission")) : V3{0, 0, 0};
      out.emission[0] = e.x; out.emission[1] = e.y; out.emission[2] = e.z;
      out.two_sided = l.get("two-sided") ? (l.get("two-sided")->as_bool() ? 1u : 0u) : 1u;
      scene->lights.push_back(out);
    }
  }
  // ---- instances (Loader.fs:118-145)
  std::vector<InstanceObj> objs;
  {
    const bnjson::Value& is = need("instances");
    for (size_t i = 0; i < is.size(); ++i) {
      const bnjson::Value& in = is[i];
      InstanceObj o;
      int p = in.get("primitive") ? in.get("primitive")->as_int() : 0;
      o.prim = prims.at(p);
      if (auto* m = in.get("material")) { o.material = m->as_int(); if (o.material < 0 || o.material >= (int)scene->materials.size()) throw std::runtime_error("material index out of range"); }
      if (auto* l = in.get("light")) { o.light = l->as_int(); if (o.light < 0 || o.light >= (int)scene->lights.size()) throw std::runtime_error("light index out of range"); }
      objs.push_back(o);
    }
  }
  // ---- nodes (Loader.fs:147-168)
  std::vector<NodeObj> nodes;
  {
    const bnjson::Value& ns = need("nodes");
    for (size_t i = 0; i < ns.size(); ++i) {
      const bnjson::Value& n = ns[i];
      NodeObj o;
      if (auto* t = n.get("transform")) {
        const TransformObj& tr = transforms.at(t->as_int());
        for (auto& k : tr.keys) { o.key_time.push_back(k.first); o.key_matrix.push_back(k.second); }
      }
      if (auto* in = n.get("instances")) for (size_t k = 0; k < in->size(); ++k) o.instances.push_back((*in)[k].as_int());
      if (auto* ch = n.get("children")) for (size_t k = 0; k < ch->size(); ++k) o.children.push_back((*ch)[k].as_int());
      o.has_camera = n.get("has-camera") ? n.get("has-camera")->as_bool() : false;
      nodes.push_back(std::move(o));
    }
  }
  // ---- integrator / camera / film (Loader.fs:170-239)
  {
    const bnjson::Value& in = need("integrator");
    const std::string& type = in.get("type") ? in.get("type")->as_string() : std::string();
    if (type == "normal") scene->info.integrator = 0;
    else if (type == "direct") scene->info.integrator = 1;
    else if (type == "path-tracing") scene->info.integrator = 2;
    else if (type == "pssmlt") scene->info.integrator = 3;
    else throw std::runtime_error("Unknown integrator type: " + type);
    scene->info.spp = in.get("spp") ? in.get("spp")->as_int() : 1;
    scene->info.max_depth = in.get("max-depth") ? in.get("max-depth")->as_int() : 8;
    scene->info.rr_depth = in.get("rr-depth") ? in.get("rr-depth")->as_int() : 5;
    scene->info.n_bootstrap = in.get("n-bootstrap") ? in.get("n-bootstrap")->as_int() : 4 * 1024 * 1024;
    scene->info.n_chains = in.get("n-chains") ? in.get("n-chains")->as_int() : 1024;
    scene->info.large_step_prob = in.get("large-step-prob") ? in.get("large-step-prob")->as_f32() : 0.5f;
    scene->info.mutation_strategy = BN_MLT_GAUSSIAN;
    if (const bnjson::Value* ms = in.get("mutation-strategy")) {
      if (ms->as_string() == "Gaussian") scene->info.mutation_strategy = BN_MLT_GAUSSIAN;
      else if (ms->as_string() == "Kelemen") scene->info.mutation_strategy = BN_MLT_KELEMEN;
      else throw std::runtime_error("Unknown mutation strategy: " + ms->as_string());
    }
  }
  BnCamera cam{};
  {
    const bnjson::Value& c = need("camera");
    const std::string& type = c.get("type") ? c.get("type")->as_string() : std::string();
    if (type == "pinhole") cam.type = BN_CAM_PINHOLE;
    else if (type == "thin-lens") cam.type = BN_CAM_THIN_LENS;
    else throw std::runtime_error("Unknown camera type: " + type);
    cam.fov_y = c.get("fov") ? c.get("fov")->as_f32() : 45.f;
    cam.aspect_ratio = c.get("aspect-ratio") ? c.get("aspect-ratio")->as_f32() : 1.f;
    cam.aperture = c.get("aperture") ? c.get("aperture")->as_f32() : 0.f;
    cam.focus_distance = c.get("focus-distance") ? c.get("focus-distance")->as_f32() : 1.f;
    cam.push_forward = c.get("push-forward") ? c.get("push-forward")->as_f32() : 0.f;
    // CameraToWorld is Unchecked.defaultof (all zeros) until a has-camera node is traversed (Camera.fs:7)
    std::memset(cam.camera_to_world, 0, sizeof cam.camera_to_world);
  }
  {
    const bnjson::Value& f = need("film");
    scene->info.width = f.get("width") ? f.get("width")->as_int() : 0;
    scene->info.height = f.get("height") ? f.get("height")->as_int() : 0;
    const std::string& tm = f.get("tone-mapping") ? f.get("tone-mapping")->as_string() : std::string();
    if (tm == "identity") scene->info.tone_mapping = 0;
    else if (tm == "aces") scene->info.tone_mapping = 1;
    else if (tm == "gamma") scene->info.tone_mapping = 2;
    else throw std::runtime_error("Unknown tone mapping type: " + tm);
  }

  // ---- Scene.Traverse(t) (Base/Scene.fs:58-61)
  std::vector<int> order;
  M4 c2w;
  std::memset(c2w.m, 0, sizeof c2w.m);
  int root_idx = root.get("root") ? root.get("root")->as_int() : 0;
  traverse(nodes, root_idx, time, M4::identity(), objs, order, c2w, scene->meshes, scene->sphere_radii, 0);
  std::memcpy(cam.camera_to_world, c2w.m, sizeof c2w.m);

  // ---- TLAS: BVHNode.Build(instances, _.Bounds) permutes the instance array
  std::vector<AABB> boxes(order.size());
  for (size_t i = 0; i < order.size(); ++i) boxes[i] = objs[order[i]].bounds;
  if (order.empty()) throw std::runtime_error("scene has no instances");
  scene->tlas = bvh_build(boxes, scene->instance_perm);
  for (size_t i = 0; i < order.size(); ++i) {
    const InstanceObj& o = objs[order[scene->instance_perm[i]]];
    BnInstance r{};
    r.prim_kind = (uint32_t)o.prim.kind;
    r.prim_id = (uint32_t)o.prim.id;
    r.material_id = o.material;
    r.light_id = o.light;
    std::memcpy(r.object_to_world, o.o2w.m, 64);
    std::memcpy(r.world_to_object, o.w2o.m, 64);
    r.bounds_min[0] = o.bounds.lo.x; r.bounds_min[1] = o.bounds.lo.y; r.bounds_min[2] = o.bounds.lo.z;
    r.bounds_max[0] = o.bounds.hi.x; r.bounds_max[1] = o.bounds.hi.y; r.bounds_max[2] = o.bounds.hi.z;
    scene->instances.push_back(r);
    if (o.light >= 0) scene->light_instances.push_back((uint32_t)i);  // Array.filter HasLight, LightSampler.fs:7
  }
  // instance_perm currently indexes `order`; report original instances[] ids
  for (auto& p : scene->instance_perm) p = (uint32_t)order[p];

  // ---- flatten meshes
  for (const MeshData& m : scene->meshes) {
    BnMesh r{};
    r.vertex_offset = (uint32_t)(scene->vertices.size() / 3);
    r.vertex_count = (uint32_t)m.vertices.size();
    r.tri_offset = (uint32_t)(scene->triangles.size() / 3);
    r.tri_count = (uint32_t)(m.tri.size() / 3);
    r.node_offset = (uint32_t)scene->blas.size();
    r.node_count = (uint32_t)m.nodes.size();
    r.alias_offset = (uint32_t)scene->alias.size();
    for (const V3& v : m.vertices) { scene->vertices.push_back(v.x); scene->vertices.push_back(v.y); scene->vertices.push_back(v.z); }
    scene->triangles.insert(scene->triangles.end(), m.tri.begin(), m.tri.end());
    scene->blas.insert(scene->blas.end(), m.nodes.begin(), m.nodes.end());
    scene->alias.insert(scene->alias.end(), m.alias.begin(), m.alias.end());
    scene->mesh_recs.push_back(r);
  }

  BnSceneDesc& d = scene->desc;
  d.tlas_nodes = scene->tlas.data(); d.tlas_node_count = (uint32_t)scene->tlas.size();
  d.instances = scene->instances.data(); d.instance_count = (uint32_t)scene->instances.size();
  d.light_instances = scene->light_instances.data(); d.light_instance_count = (uint32_t)scene->light_instances.size();
  d.meshes = scene->mesh_recs.data(); d.mesh_count = (uint32_t)scene->mesh_recs.size();
  d.vertices = scene->vertices.data(); d.vertex_count = (uint32_t)(scene->vertices.size() / 3);
  d.triangles = scene->triangles.data(); d.triangle_count = (uint32_t)(scene->triangles.size() / 3);
  d.blas_nodes = scene->blas.data(); d.blas_node_count = (uint32_t)scene->blas.size();
  d.alias = scene->alias.data(); d.alias_count = (uint32_t)scene->alias.size();
  d.sphere_radii = scene->sphere_radii.data(); d.sphere_count = (uint32_t)scene->sphere_radii.size();
  d.materials = scene->materials.data(); d.material_count = (uint32_t)scene->materials.size();
  d.lights = scene->lights.data(); d.light_count = (uint32_t)scene->lights.size();
  d.camera = cam;
  return scene.release();
}

int guarded_load(const std::string& text, const char* base_dir, float time, BnHostScene** out) {
  if (!out) { set_error("bn_host_scene_load: out is NULL"); return BN_ERR_INVALID; }
  *out = nullptr;
  try {
    *out = build_scene(text, base_dir ? base_dir : "", time);
    return BN_OK;
  } catch (const std::exception& e) {
    set_error(e.what());
    return BN_ERR_IO;
  }
}

}  // namespace

extern "C" {

int bn_host_scene_load_string(const char* json_text, const char* base_dir, float time, BnHostScene** out) {
  if (!json_text) { set_error("bn_host_scene_load_string: json_text is NULL"); return BN_ERR_INVALID; }
  return guarded_load(json_text, base_dir, time, out);
}

int bn_host_scene_load(const char* json_path, const char* base_dir, float time, BnHostScene** out) {
  if (!json_path) { set_error("bn_host_scene_load: json_path is NULL"); return BN_ERR_INVALID; }
  std::ifstream in(json_path, std::ios::binary);
  if (!in) { set_error(std::string("cannot open scene file: ") + json_path); return BN_ERR_IO; }
  std::stringstream ss;
  ss << in.rdbuf();
  return guarded_load(ss.str(), base_dir, time, out);
}

int bn_host_scene_load_ex(const char* json_path, const char* base_dir, float time, int build_device, BnHostScene** out) {
  struct Restore { int prev; ~Restore() { g_build_device = prev; } } restore{g_build_device};
  g_build_device = build_device;
  return bn_host_scene_load(json_path, base_dir, time, out);
}

const BnSceneDesc* bn_host_scene_desc(const BnHostScene* s) { return s ? &s->desc : nullptr; }
void bn_host_scene_info(const BnHostScene* s, BnHostSceneInfo* info) { if (s && info) *info = s->info; }
const uint32_t* bn_host_scene_instance_permutation(const BnHostScene* s) { return s ? s->instance_perm.data() : nullptr; }
const uint32_t* bn_host_scene_triangle_permutation(const BnHostScene* s, uint32_t mesh) {
  return (s && mesh < s->meshes.size()) ? s->meshes[mesh].perm.data() : nullptr;
}
void bn_host_scene_destroy(BnHostScene* s) { delete s; }

int bn_host_bvh_build(const float* boxes, uint32_t n, BnBVHNode* nodes, uint32_t max_nodes, uint32_t* perm) {
  if (!boxes || !nodes || n == 0) { set_error("bn_host_bvh_build: bad arguments"); return BN_ERR_INVALID; }
  try {
    std::vector<AABB> b(n);
    for (uint32_t i = 0; i < n; ++i) b[i] = {{boxes[i * 6], boxes[i * 6 + 1], boxes[i * 6 + 2]}, {boxes[i * 6 + 3], boxes[i * 6 + 4], boxes[i * 6 + 5]}};
    std::vector<uint32_t> p;
    auto out = bvh_build(b, p);
    if (out.size() > max_nodes) { set_error("bn_host_bvh_build: node buffer too small"); return BN_ERR_INVALID; }
    std::memcpy(nodes, out.data(), out.size() * sizeof(BnBVHNode));
    if (perm) std::memcpy(perm, p.data(), p.size() * sizeof(uint32_t));
    return (int)out.size();
  } catch (const std::exception& e) {
    set_error(e.what());
    return BN_ERR_INVALID;
  }
}

int bn_host_alias_build(const float* weights, uint32_t n, BnAliasEntry* out) {
  if (!weights || !out) { set_error("bn_host_alias_build: bad arguments"); return BN_ERR_INVALID; }
  std::vector<float> w(weights, weights + n);
  auto t = alias_build(w);
  std::memcpy(out, t.data(), t.size() * sizeof(BnAliasEntry));
  return BN_OK;
}

// Film.PostProcess (Base/Film.fs:21-30) + Rgba32(Vector3) conversion (Film.fs:64):
// ImageSharp's Rgba32(Vector3) clamps to [0,1], scales by 255 and rounds to
// nearest (MidpointRounding.AwayFromZero in its Pack()).
int bn_host_film_to_rgba8(const float* film, int32_t w, int32_t h, int32_t tone, uint8_t* rgba) {
  if (!film || !rgba || w <= 0 || h <= 0) { set_error("bn_host_film_to_rgba8: bad arguments"); return BN_ERR_INVALID; }
  const float gamma = 1.f / 2.2f;
  for (int64_t i = 0; i < (int64_t)w * h; ++i) {
    for (int c = 0; c < 3; ++c) {
      float x = film[i * 3 + c];
      if (tone == 1) x = x * (2.51f * x + 0.03f) / (x * (2.43f * x + 0.59f) + 0.14f);
      else if (tone == 2) x = powf(x, gamma);
      // Vector3.Clamp = Min(Max(x, 0), 1); NaN -> 0 after the byte conversion
      x = !(x > 0.f) ? 0.f : (x > 1.f ? 1.f : x);
      rgba[i * 4 + c] = (uint8_t)(x * 255.f + 0.5f);
    }
    rgba[i * 4 + 3] = 255;
  }
  return BN_OK;
}

}  // extern "C"
