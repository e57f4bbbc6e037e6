// raydar-cuda -- headless driver over the C ABI, mirroring the reference's `raydar` binary
// (src/main.rs:10-107, flags of src/cli/mod.rs:12-27,66-68):
//
//   raydar-cuda [--max-sample-count N] [--max-bounces B] [-o out.png] [--gpus G] [--partition samples|stripes] [--resolution WxH]
//               [--seed S] [scene.rscn]
//
// Same info / profiling printout as the reference plus a samples/s line.  No scene file = Scene::default()
// (cli/mod.rs:38-40).  --resolution recomputes the camera matrices (Camera::set_resolution_x/y); the reference has
// no such flag (resolution comes from the scene file only).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "raydar_cuda.h"

static int die(const char *what, RdrRenderer *r)
{
    fprintf(stderr, "Error: %s: %s\n", what, rdr_last_error(r));
    return 1;
}

int main(int argc, char **argv)
{
    RdrConfig config{1024u, 12u};                         // RendererConfig::default(), renderer/mod.rs:16-23
    std::string output = "output.png", scene_file;
    int gpus = 1, partition = RDR_PARTITION_SAMPLES;
    uint32_t width = 0, height = 0;
    unsigned long long seed = 0x5EED;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&](const char *name) -> const char * {
            if (i + 1 >= argc) { fprintf(stderr, "error: a value is required for '%s'\n", name); exit(2); }
            return argv[++i];
        };
        if (a == "--max-sample-count") config.max_sample_count = (uint32_t)strtoul(value("--max-sample-count"), nullptr, 10);
        else if (a == "--max-bounces") config.max_bounces = (uint32_t)strtoul(value("--max-bounces"), nullptr, 10);
        else if (a == "-o" || a == "--output") output = value("--output");
        else if (a == "--gpus") gpus = atoi(value("--gpus"));
        else if (a == "--partition") {
            const std::string v = value("--partition");
            if (v == "samples") partition = RDR_PARTITION_SAMPLES;
            else if (v == "stripes") partition = RDR_PARTITION_STRIPES;
            else { fprintf(stderr, "error: --partition expects samples|stripes\n"); return 2; }
        }
        else if (a == "--seed") seed = strtoull(value("--seed"), nullptr, 0);
        else if (a == "--resolution") {
            if (sscanf(value("--resolution"), "%ux%u", &width, &height) != 2) { fprintf(stderr, "error: --resolution expects WxH\n"); return 2; }
        } else if (a == "--cuda" || a == "--cpu") {
            if (a == "--cpu") { fprintf(stderr, "error: this binary only has the CUDA backend (no CPU fallback)\n"); return 2; }
        } else if (a == "-h" || a == "--help") {
            printf("Usage: raydar-cuda [--max-sample-count N] [--max-bounces B] [-o out.png] [--gpus G] [--partition samples|stripes] [--resolution WxH] [--seed S] [scene.rscn]\n");
            return 0;
        } else if (!a.empty() && a[0] == '-') { fprintf(stderr, "error: unexpected argument '%s'\n", a.c_str()); return 2; }
        else scene_file = a;
    }

    RdrScene *scene = nullptr;
    if (!scene_file.empty() ? rdr_scene_load_rscn(scene_file.c_str(), &scene) : rdr_scene_default(&scene)) return die("scene", nullptr);
    if (width && height && rdr_scene_set_resolution(scene, width, height)) return die("resolution", nullptr);
    RdrSceneFlat flat;
    rdr_scene_flat(scene, &flat);

    RdrRenderer *r = nullptr;
    std::vector<int> devices;
    for (int g = 0; g < gpus; ++g) devices.push_back(g);
    if ((gpus > 1 ? rdr_create_multi(&config, gpus, devices.data(), &r) : rdr_create(&config, 0, &r)) != RDR_OK) return die("renderer", nullptr);
    rdr_set_seed(r, seed);
    if (gpus > 1 && rdr_set_partition(r, partition, 0) != RDR_OK) return die("partition", r);

    // print_info, main.rs:26-59
    printf("=== Raydar (CUDA backend) %s ===\n", rdr_version());
    printf("Renderer: CUDA x%d\n", gpus);
    printf("Max Samples: %u\n", rdr_max_sample_count(r));
    printf("Max Bounces: %u\n", rdr_max_bounces(r));
    printf("Resolution: %ux%u\n", flat.width, flat.height);
    printf("Objects: %u\n", flat.n_objects);

    std::vector<uint8_t> image((size_t)flat.width * flat.height * 4);
    if (rdr_render_frame(r, &flat, image.data()) != RDR_OK) return die("render_frame", r);
    if (rdr_write_png(output.c_str(), image.data(), flat.width, flat.height) != RDR_OK) return die("Cannot save image", nullptr);

    // print_profiling_metrics, main.rs:61-107 (an unset timer is an error there too)
    RdrProfiler p;
    rdr_profiler(r, &p);
    printf("\n=== Render Profiling Metrics ===\n");
    if (!p.has_prepare) { fprintf(stderr, "Error: Prepare timer not started\n"); return 1; }
    printf("Scene Preparation: %llums\n", (unsigned long long)(p.prepare_ns / 1000000));
    if (!p.has_render) { fprintf(stderr, "Error: Render timer not started\n"); return 1; }
    printf("Render Time: %llums\n", (unsigned long long)(p.render_ns / 1000000));
    if (!p.has_sample) { fprintf(stderr, "Error: Sample timer not started\n"); return 1; }
    printf("Last Sample Time: %llums\n", (unsigned long long)(p.sample_ns / 1000000));
    printf("----------------------------------------\n");
    if (!p.has_frame) { fprintf(stderr, "Error: Frame timer not started\n"); return 1; }
    printf("Total Frame Time: %llums\n", (unsigned long long)(p.frame_ns / 1000000));
    const double samples = (double)flat.width * flat.height * rdr_sample_count(r);
    printf("Throughput: %.1f Msamples/s (kernel %.1f ms)\n", samples / (p.render_ns * 1e-9) / 1e6, p.device_render_ms);

    rdr_destroy(r);
    rdr_scene_free(scene);
    return 0;
}
