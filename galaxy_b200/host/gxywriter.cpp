// gxywriter -- state file in, PNG images out: the batch renderer of the reference (src/apps/gxywriter.cpp:61-291)
// on top of the B200 C ABI.  Same command line where it applies, same loop order (for every camera, for every
// visualization), same image names (Rendering::SaveImage), same "TIMING total" line.
//
//   gxywriter [-s width height] [-S skip] [-P partitions] [-d device] [-o basename] [--describe] statefile
//
//   -s w h      override the camera window (reference: -s; default: the camera's own 512x512, Camera.h:203-204)
//   -S k        render only every k-th (camera, visualization) pair (reference: -S)
//   -P n        spatial partitions (the reference takes this from the MPI size; here the partitions live on
//               one device and exchange rays by device copies; one process per GPU goes through gxy_comm_init)
//   --describe  parse everything, print the parsed state as JSON and exit without touching a GPU (tests)
// Not provided: -C (Cinema database), -c (client/server), -A/-D (debugger attach), -N.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "gxy_host.h"

static void syntax(const char *a) {
  std::cerr << "syntax: " << a << " [options] statefile\n"
            << "options:\n"
            << "  -s w h      window width, height (camera default 512 x 512)\n"
            << "  -S k        only render every k'th rendering\n"
            << "  -P n        number of spatial partitions (1)\n"
            << "  -d dev      CUDA device (0)\n"
            << "  -o base     image base name (image)\n"
            << "  --describe  print the parsed state as JSON and exit (no GPU needed)\n";
  exit(1);
}

int main(int argc, char *argv[]) {
  std::string statefile, base = "image";
  int width = 1920, height = 1080, skip = 0, nparts = 1, device = 0;
  bool override_windowsize = false, describe = false;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-s") && i + 2 < argc) { width = atoi(argv[++i]); height = atoi(argv[++i]); override_windowsize = true; }
    else if (!strcmp(argv[i], "-S") && i + 1 < argc) skip = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-P") && i + 1 < argc) nparts = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-d") && i + 1 < argc) device = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-o") && i + 1 < argc) base = argv[++i];
    else if (!strcmp(argv[i], "--describe")) describe = true;
    else if (argv[i][0] != '-' && statefile.empty()) statefile = argv[i];
    else syntax(argv[0]);
  }
  if (statefile.empty() || nparts < 1) syntax(argv[0]);

  gxy::json::Value doc;
  try {
    doc = gxy::json::ParseFile(statefile);
  } catch (const std::exception &e) {
    std::cerr << "Bad state file: " << statefile << " (" << e.what() << ")\n";
    return 1;
  }
  const size_t slash = statefile.find_last_of('/');
  const std::string state_dir = slash == std::string::npos ? std::string("") : statefile.substr(0, slash + 1);

  gxy::Renderer theRenderer;
  theRenderer.LoadStateFromDocument(doc);
  std::vector<gxy::Camera> theCameras;
  if (!gxy::Camera::LoadCamerasFromJSON(doc, theCameras, state_dir)) { std::cerr << "error loading cameras\n"; return 1; }
  gxy::Datasets theDatasets;
  if (!theDatasets.LoadFromJSON(doc, state_dir)) { std::cerr << "error loading theDatasets\n"; return 1; }
  std::vector<gxy::Visualization> theVisualizations;
  if (!gxy::Visualization::LoadVisualizationsFromJSON(doc, theVisualizations, state_dir)) { std::cerr << "error loading visualizations\n"; return 1; }

  if (describe) {
    std::cout << gxy::describe_state(theRenderer, theCameras, theVisualizations, theDatasets, nparts) << std::endl;
    return 0;
  }

  if (gxy_device_count() <= 0) {
    std::cerr << "no CUDA device: " << gxy_last_error() << " (galaxy_b200 has no CPU fallback)\n";
    return 1;
  }
  gxy_context *ctx = nullptr;
  if (gxy_context_create(device, &ctx)) { std::cerr << "gxy_context_create: " << gxy_last_error() << "\n"; return 1; }
  for (auto &v : theVisualizations)
    if (!v.Commit(ctx, theDatasets, nparts)) { std::cerr << "error committing a visualization\n"; return 1; }

  const auto t_rendering_start = std::chrono::steady_clock::now();
  int k = 0, index = 0;
  long long rays = 0;
  std::cout << "render start" << std::endl;
  for (auto &c : theCameras)
    for (auto &v : theVisualizations) {
      if (skip && (k % skip) != 0) { std::cerr << "S"; k++; continue; }
      // a sampling Visualization (only GradientSampler / IsoSampler operators, src/sampler): there is nothing to render; the rays
      // of this camera leave samples, written as <base>_%05d.samples (raw float32 xyz, partition after partition).  Not part of
      // the reference's gxywriter, which cannot load such a Visualization at all: the Sampler is driven from its GUI server.
      bool sampling = !v.operators.empty();
      for (const auto &op : v.operators) sampling = sampling && (op.type == "GradientSamplerVis" || op.type == "IsoSamplerVis");
      if (sampling) {
        gxy::Sampler theSampler;
        if (!theSampler.Sample(c, v, override_windowsize ? width : c.width, override_windowsize ? height : c.height)) { std::cerr << "error sampling\n"; return 1; }
        char name[32];
        snprintf(name, sizeof name, "_%05d.samples", index);
        FILE *f = fopen((base + name).c_str(), "wb");
        if (!f) { std::cerr << "error saving " << base << name << "\n"; return 1; }
        long long total = 0;
        for (size_t r = 0; r < v.parts.size(); r++) {
          std::vector<float> xyz;
          if (!theSampler.GetSamples(v, (int)r, xyz)) { fclose(f); return 1; }
          if (!xyz.empty()) fwrite(xyz.data(), sizeof(float), xyz.size(), f);
          total += (long long)(xyz.size() / 3);
        }
        fclose(f);
        std::cout << "samples = " << total << std::endl;
        rays += theSampler.stats.traced_rays;
        index++;
        k++;
        continue;
      }
      gxy::Rendering theRendering;
      theRendering.camera = &c;
      theRendering.visualization = &v;
      theRendering.width = override_windowsize ? width : c.width;
      theRendering.height = override_windowsize ? height : c.height;
      if (!theRendering.Render(theRenderer)) { std::cerr << "error rendering\n"; return 1; }
      if (!theRendering.SaveImage(base, index)) { std::cerr << "error saving " << theRendering.ImageName(base, index) << "\n"; return 1; }
      rays += theRendering.stats.primary_rays + theRendering.stats.shadow_rays + theRendering.stats.ao_rays;
      index++;
      k++;
    }
  std::cout << "index = " << index << std::endl;
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_rendering_start).count();
  std::cout << index << ": " << secs << " seconds (" << rays << " rays)" << std::endl;
  std::cout << "TIMING total " << secs << " seconds" << std::endl;
  for (auto &v : theVisualizations) v.Release();
  gxy_context_destroy(ctx);
  std::cerr << "done\n";
  return 0;
}
