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
  bn::GCamera cam;
  int max_stack = 0;
  bool all_finite = true;
};

// Validates the description (indices in range, well-formed preorder trees, leaf
// sizes and depths the traversal stack can hold) and fills `out`.
bool convert_scene(const BnSceneDesc& d, ConvertedScene& out, std::string& err);

}  // namespace bnconv
