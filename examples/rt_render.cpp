// rt_render — a host program written only against the C headers (include/*.h): builds one of the named scenes,
// renders N frames through the per-frame orchestrator and writes the tonemapped result as a PNG. It is the C++
// shape of what the reference's Renderer + FramePresenter do per frame (Renderer.swift:1405-1503) and what a port of
// the app would look like; the Python package is just another client of the same two libraries.
//
//   rt_render <scene> <width> <height> <spp> <maxBounces> <frames> <out.png> [assetDir|-] [environment.hdr] [intensity]
//
// With a Radiance .hdr file the environment extension is bound (rt_b200.h rt_environment) and sampled as a light
// (RT_ENV_IMPORTANCE); without one a ray that leaves the scene returns black, as in the reference.
//
// Exit code 0 and one line "frames=.. ms_per_frame=.. mrays_per_s=.." on success; errors come back as messages from
// rt_last_error / rtr_last_error / rts_last_error (the library never aborts).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/rt_b200.h"
#include "../include/rt_renderer.h"
#include "../include/rt_scene.h"

#define CHECK(expr, who)                                          \
  do {                                                            \
    if ((expr) != 0) {                                            \
      std::fprintf(stderr, "%s failed: %s\n", #expr, who());      \
      return 1;                                                   \
    }                                                             \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 8) {
    std::fprintf(stderr, "usage: %s <scene> <width> <height> <spp> <maxBounces> <frames> <out.png> [assetDir|-] [environment.hdr] [intensity]\n", argv[0]);
    return 2;
  }
  const char *name = argv[1];
  const int width = std::atoi(argv[2]), height = std::atoi(argv[3]), spp = std::atoi(argv[4]), bounces = std::atoi(argv[5]);
  const int frames = std::atoi(argv[6]);
  const char *outPath = argv[7];
  const char *assetDir = (argc > 8 && std::strcmp(argv[8], "-") != 0) ? argv[8] : nullptr;
  const char *envPath = argc > 9 ? argv[9] : nullptr;
  const float envIntensity = argc > 10 ? float(std::atof(argv[10])) : 1.0f;

  rt_uniforms uniforms{};
  uint32_t seed = 0;
  rts_scene *scene = rts_scene_create_named(name, assetDir, width, height, &uniforms, &seed);
  if (!scene) {
    std::fprintf(stderr, "rts_scene_create_named(%s) failed: %s\n", name, rts_last_error());
    return 1;
  }
  uniforms.samplesPerPixel = spp;
  uniforms.maxBounces = bounces;
  rt_scene_desc desc{};
  CHECK(rts_scene_get_desc(scene, &desc), rts_last_error);

  rt_context *ctx = nullptr;
  CHECK(rt_create(0, &ctx), rt_last_error);
  rtr_renderer *renderer = nullptr;
  CHECK(rtr_create(ctx, &desc, width, height, 0, &renderer), rtr_last_error);
  std::vector<uint32_t> seeds(size_t(width) * height);
  rts_fill_seed_image(seeds.data(), width, height, seed);
  CHECK(rtr_set_seeds(renderer, seeds.data()), rtr_last_error);

  uint64_t *countersDev = nullptr;
  CHECK(rt_malloc(ctx, 192, reinterpret_cast<void **>(&countersDev)), rt_last_error);
  CHECK(rt_memset(ctx, countersDev, 0, 192), rt_last_error);
  rt_trace_options options{};
  options.rayCountersDev = countersDev;

  // environment extension: texels + importance-sampling table, both built on the host and uploaded once
  rt_environment environment{};
  float *envTexelsDev = nullptr, *envCdfDev = nullptr;
  if (envPath) {
    int ew = 0, eh = 0;
    float *texels = nullptr;
    CHECK(rts_load_hdr(envPath, &ew, &eh, &texels), rts_last_error);
    std::vector<float> cdf(rt_environment_cdf_floats(ew, eh));
    if (rt_environment_cdf(texels, ew, eh, cdf.data()) != 0) {
      std::fprintf(stderr, "rt_environment_cdf failed\n");
      return 1;
    }
    const size_t texelBytes = size_t(ew) * size_t(eh) * 4 * sizeof(float);
    CHECK(rt_malloc(ctx, texelBytes, reinterpret_cast<void **>(&envTexelsDev)), rt_last_error);
    CHECK(rt_malloc(ctx, cdf.size() * sizeof(float), reinterpret_cast<void **>(&envCdfDev)), rt_last_error);
    CHECK(rt_upload(ctx, envTexelsDev, texels, texelBytes), rt_last_error);
    CHECK(rt_upload(ctx, envCdfDev, cdf.data(), cdf.size() * sizeof(float)), rt_last_error);
    CHECK(rt_sync(ctx), rt_last_error); // the uploads read host memory that goes away below
    rts_free(texels);
    environment.texelsDev = envTexelsDev;
    environment.width = ew, environment.height = eh;
    environment.intensity = envIntensity;
    environment.flags = RT_ENV_IMPORTANCE | RT_ENV_GUIDED; // the table below comes from rt_environment_cdf
    environment.cdfDev = envCdfDev;
    options.environment = &environment;
  }

  CHECK(rt_timer_begin(ctx), rt_last_error);
  for (int f = 0; f < frames; ++f) {
    uniforms.frameIndex = uint32_t(f);
    if (f > 0) { // Renderer.updateSkinningAndBLAS: animate, re-skin, refit, rebuild the TLAS
      CHECK(rts_animate(scene, f / 60.0), rts_last_error);
      CHECK(rts_scene_get_desc(scene, &desc), rts_last_error);
      CHECK(rtr_update(renderer, &desc), rtr_last_error);
    }
    CHECK(rtr_draw(renderer, &uniforms, &options), rtr_last_error);
  }
  float ms = 0.0f;
  CHECK(rt_timer_end(ctx, &ms), rt_last_error);
  uint64_t counters[3] = {0, 0, 0};
  CHECK(rt_download(ctx, counters, countersDev, sizeof counters), rt_last_error);

  // display transform + image writer (Shaders.metal:38-52): Reinhard, sRGB, row 0 = top
  rt_image frame{};
  CHECK(rtr_image_info(renderer, RT_TEXTURE_ACCUMULATION, &frame), rtr_last_error);
  uint8_t *rgbaDev = nullptr;
  CHECK(rt_malloc(ctx, size_t(width) * height * 4, reinterpret_cast<void **>(&rgbaDev)), rt_last_error);
  CHECK(rt_tonemap(ctx, &frame, rgbaDev, RT_TONEMAP_SRGB | RT_TONEMAP_FLIP_Y), rt_last_error);
  std::vector<uint8_t> rgba(size_t(width) * height * 4);
  CHECK(rt_download(ctx, rgba.data(), rgbaDev, rgba.size()), rt_last_error);
  CHECK(rts_write_png(outPath, rgba.data(), width, height), rts_last_error);

  const double rays = double(counters[0] + counters[1]);
  std::printf("scene=%s %dx%d spp=%d bounces=%d frames=%d ms_per_frame=%.3f mrays_per_s=%.1f launches=%llu png=%s\n", name,
              width, height, spp, bounces, frames, ms / frames, rays / (ms * 1e-3) / 1e6,
              (unsigned long long)rt_launch_count(ctx), outPath);
  rt_free(ctx, rgbaDev);
  rt_free(ctx, countersDev);
  if (envTexelsDev) rt_free(ctx, envTexelsDev);
  if (envCdfDev) rt_free(ctx, envCdfDev);
  rtr_destroy(renderer);
  rt_destroy(ctx);
  rts_scene_destroy(scene);
  return 0;
}
