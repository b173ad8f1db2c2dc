// barnacle_gpu — C++ caller of the C ABI that mirrors the reference's Program.fs
// (`-i scene.json -o image`, Program.fs:7-27) and Scene.Render (Render.fs:10-19):
// load, render on the GPU, print the Stopwatch line, save.  Output formats by
// extension: .ppm (tone-mapped RGB8, what Film.Save encodes) or .pfm (linear fp32).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/barnacle_b200.h"

static int fail(const char* what) {
  std::fprintf(stderr, "%s: %s\n", what, bn_last_error());
  return 1;
}

int main(int argc, char** argv) {
  std::string input, output, base_dir;
  int device = 0;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if ((a == "-i" || a == "--input") && i + 1 < argc) input = argv[++i];
    else if ((a == "-o" || a == "--output") && i + 1 < argc) output = argv[++i];
    else if (a == "--base-dir" && i + 1 < argc) base_dir = argv[++i];
    else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
  }
  if (input.empty() || output.empty()) {
    std::fprintf(stderr, "Invalid arguments.\nusage: barnacle_gpu -i <scene.json> -o <image.ppm|image.pfm> [--base-dir DIR] [--device N]\n");
    return 2;
  }
  BnHostScene* host = nullptr;
  if (bn_host_scene_load(input.c_str(), base_dir.empty() ? nullptr : base_dir.c_str(), 0.f, &host) != BN_OK) return fail("Scene.Load");
  std::printf("Loaded scene from %s\n", input.c_str());
  BnHostSceneInfo info{};
  bn_host_scene_info(host, &info);
  BnScene* scene = nullptr;
  if (bn_scene_create(bn_host_scene_desc(host), device, &scene) != BN_OK) return fail("bn_scene_create");
  std::vector<float> film((size_t)info.width * info.height * 3, 0.f);  // Film.Clear (Render.fs:11)
  // Loader.fs:185-204 order: 0 normal, 1 direct, 2 path-tracing, 3 pssmlt
  if (info.integrator == 3) {
    BnMltParams p{};
    p.width = info.width; p.height = info.height; p.mutations_per_pixel = info.spp; p.max_depth = info.max_depth; p.rr_depth = info.rr_depth;
    p.n_bootstrap = info.n_bootstrap; p.n_chains = info.n_chains; p.strategy = info.mutation_strategy;
    if (p.strategy == BN_MLT_KELEMEN) { p.p0 = 1.f / 1024.f; p.p1 = 1.f / 16.f; }  // MutationStrategy.Kelemen(1/1024, 1/16), Loader.fs:199
    else { p.p0 = 1e-2f; }                                                          // MutationStrategy.Gaussian(1e-2), Loader.fs:195,198
    p.large_step_prob = info.large_step_prob;
    p.chain_begin = 0; p.chain_end = info.n_chains;
    BnMltStats st{};
    auto t0 = std::chrono::steady_clock::now();
    if (bn_render_pssmlt(scene, &p, film.data(), &st) != BN_OK) return fail("bn_render_pssmlt");
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (st.b == 0.f) {
      std::printf("Warning: all bootstrap samples are zero, exiting...\n");  // PSSMLT.fs:396-397
    } else {  // PSSMLT.fs:412-414
      std::printf("Accepted mutation count: %llu\n", (unsigned long long)st.accepted);
      std::printf("Proposed mutation count: %llu\n", (unsigned long long)st.proposed);
      std::printf("Acceptance rate: %f\n", (double)st.accepted / (double)(st.proposed ? st.proposed : 1));
    }
    std::printf("Render time: %f seconds\n", sec);
  } else {
    BnRenderParams p{};
    p.width = info.width; p.height = info.height; p.spp = info.spp; p.max_depth = info.max_depth; p.rr_depth = info.rr_depth;
    p.sample_begin = 0; p.sample_end = info.spp; p.x0 = 0; p.y0 = 0; p.x1 = info.width; p.y1 = info.height;
    p.interleave_count = 1;
    p.integrator = info.integrator == 2 ? BN_INTEGRATOR_PATH_TRACING : (info.integrator == 1 ? BN_INTEGRATOR_DIRECT : BN_INTEGRATOR_NORMAL);
    BnStats st{};
    auto t0 = std::chrono::steady_clock::now();
    if (bn_render(scene, &p, film.data(), &st) != BN_OK) return fail("bn_render");
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("Render time: %f seconds\n", sec);
    std::printf("paths %llu, rays %llu (%.1f Mrays/s on the device)\n", (unsigned long long)st.paths,
                (unsigned long long)(st.extend_rays + st.shadow_rays), (st.extend_rays + st.shadow_rays) / (st.gpu_ms * 1e3));
  }
  FILE* f = std::fopen(output.c_str(), "wb");
  if (!f) { std::fprintf(stderr, "cannot write %s\n", output.c_str()); return 4; }
  if (output.size() > 4 && output.substr(output.size() - 4) == ".pfm") {
    std::fprintf(f, "PF\n%d %d\n-1.0\n", info.width, info.height);
    for (int y = info.height - 1; y >= 0; --y) std::fwrite(&film[(size_t)y * info.width * 3], sizeof(float), (size_t)info.width * 3, f);
  } else {
    std::vector<uint8_t> rgba((size_t)info.width * info.height * 4);
    bn_host_film_to_rgba8(film.data(), info.width, info.height, info.tone_mapping, rgba.data());
    std::fprintf(f, "P6\n%d %d\n255\n", info.width, info.height);
    for (size_t i = 0; i < (size_t)info.width * info.height; ++i) std::fwrite(&rgba[i * 4], 1, 3, f);
  }
  std::fclose(f);
  bn_scene_destroy(scene);
  bn_host_scene_destroy(host);
  return 0;
}
