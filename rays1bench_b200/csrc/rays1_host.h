// rays1_host.h -- the reference's host surface on top of the C ABI (include/rays1_b200.h).
// Same names, argument meaning and ownership as src/latest/rayweek1.cpp + src/common/common.h of the reference:
//   Scene* create_small_scene() / create_medium_scene() / create_large_scene()   (rayweek1.cpp:552, 582, 654)
//   RESULT benchmark(Scene* scene, Pix* pixels, bool write_tga, const char* scene_name)  (rayweek1.cpp:845)
//   log_results(), tga_write_rgb24(), RESULT, Pix                                 (common.h:36-122)
// What the reference fixes with macros (SCREEN_W/H, NUM_SAMPLES_PER_PIXEL, MAX_BOUNCES; common.h:3-31) is a runtime
// HostConfig here, with the reference's values as defaults.
#pragma once
#include <stdint.h>

#include "../../include/rays1_b200.h"

struct HostConfig {
    int width = 1280, height = 720;  // common.h:19-20
    int spp = 10 * 25;               // common.h:25 (MULTITHREADED)
    int max_bounces = 50;            // common.h:18
    int variant = R1_VARIANT_MEGAKERNEL;
    int n_gpus = 1;
    uint32_t seed = 0;
    int row_tile = 1;                // rows per interleaved tile of the multi-GPU partition
    bool quiet = false;              // suppress the stdout report (used by ctypes callers that print their own)
};
HostConfig &host_config();

// common.h:36-45, field for field
struct RESULT {
    double elapsed_seconds;
    uint64_t num_rays;

    double get_mrays_per_sec() const { return elapsed_seconds ? (num_rays / elapsed_seconds / 1000000.0) : 0; }
};

// common.h:80-83
struct Pix {
    uint8_t r, g, b;
};

// rayweek1.cpp:539-549 -- owns the scene; here the r1_scene handle (host SoA + device buffers).
class Scene {
public:
    r1_scene *handle = nullptr;
    ~Scene();
};

Scene *create_small_scene();
Scene *create_medium_scene();
Scene *create_large_scene();
Scene *create_synth4096_scene();  // SURVEY.md 8d config 5; not in the reference
Scene *create_scene_by_name(const char *name);
Scene *create_scene_from_file(const char *path);  // text scene description (rays1_host.cpp); nullptr on error

// rayweek1.cpp:845 -- `pixels` must hold host_config().width * height entries.  A CUDA / NCCL failure is fatal here (the
// reference has no error path); the C-linkage form r1_host_benchmark (include/rays1_b200.h) returns an error code instead.
RESULT benchmark(Scene *scene, Pix *pixels, bool write_tga, const char *scene_name);
// CUDA-event time of the kernels of the last benchmark() call (what the reference's RESULT has no field for)
double benchmark_last_kernel_ms();
void log_results(const char *version, const char *scene, const RESULT *results, int num_runs);
bool tga_write_rgb24(const char *filename, int width, int height, Pix *pixels);  // !!! swaps R and B in `pixels`

// with RAYS1_REFERENCE_MAIN defined before this header, the reference's own main()
// (src/latest/rayweek1.cpp:930-988) compiles UNCHANGED against it -- the two macros it reads from common.h:19-20 map to the
// runtime configuration, and the libc headers rayweek1.cpp includes at its top are included here.
// (The test suite builds and runs exactly that: INTEGRATION.md section 2.)
#ifdef RAYS1_REFERENCE_MAIN
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define SCREEN_W (host_config().width)
#define SCREEN_H (host_config().height)
#endif
