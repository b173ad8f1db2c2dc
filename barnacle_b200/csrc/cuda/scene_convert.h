// BnSceneDesc (reference layouts) -> device layout (device_scene.h), on the host.
#pragma once
#include <string>
#include <vector>

#include "../../../include/barnacle_b200.h"
#include "device_scene.h"

namespace bnconv {

struct ConvertedScene {
  std::vector<bn::GNode> nodes;
  bn::GTree tlas;
  std::vector<bn::GInstTrav> inst_trav;
  std::vector<bn::GInstHead> inst_head;
  std::vector<bn::GMat43> inst_w2o, inst_o2w;
  std::vector<bn::GMesh> meshes;
  std::vector<bn::GTri> tris;
  std::vector<bn::GAlias> alias;
  std::vector<float> sphere_radii;
  std::vector<bn::GMaterial> materials;
  std::vector<bn::GLight> lights;
  std::vector<uint32_t> light_inst;
  std::vector<bn::GFlatInst> flat_tlas;  // 8 * n_inst entries, empty when the TLAS is not small
  // 4-wide nodes of the fast path (empty when some node of the scene fails the containment / finiteness check: the
  // binary fast path is used then) and what depends on the choice
  std::vector<bn::GWide> wide;
  uint32_t tlas_wroot = 0;
  std::vector<uint32_t> inst_wroot;      // per instance: BLAS root ref in `wide` (copied into inst_trav[].wroot / flat_tlas[].direct_root)
  int max_stack_wide = 0;
  bn::GCamera cam;
  int max_stack = 0;
  bool all_finite = true;
};

// Validates the description (indices in range, well-formed preorder trees, leaf
// sizes and depths the traversal stack can hold) and fills `out`.
bool convert_scene(const BnSceneDesc& d, ConvertedScene& out, std::string& err);
void use_binary_nodes(ConvertedScene& cs);
void convert_camera(const BnCamera& c, bn::GCamera& cam);  // the camera is applied per device scene (it is not part of the cached layout)

}  // namespace bnconv
