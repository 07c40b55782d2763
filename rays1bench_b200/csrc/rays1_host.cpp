// rays1_host.cpp -- the reference's host surface (scene builders, benchmark(), report, TGA, log) on the C ABI.
// Mirrors src/latest/rayweek1.cpp:552-719, 845-927 and src/common/common.h:36-122 of the reference; the trace loop
// itself runs on the GPU(s) behind r1_render / r1_render_device.  file:line citations are relative to /root/reference/.
#include "rays1_host.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

HostConfig &host_config()
{
    static HostConfig cfg;
    return cfg;
}

Scene::~Scene()
{
    // Hitable::~Hitable deletes the materials and the SoA buffer (rayweek1.cpp:141-149); here: host SoA + device buffers.
    r1_scene_destroy(handle);
}

namespace {

[[noreturn]] void die(const char *what)
{
    // the reference has no error paths (-fno-exceptions, bench.py:175); a CUDA/NCCL failure is fatal, never a CPU fallback
    fprintf(stderr, "rays1_b200: %s: %s\n", what, r1_last_error());
    exit(1);
}

// Scene builders.  `commit` = also emit the device buffers (one replica per GPU that will render); the C++ surface
// always commits, ctypes callers may build host-only scenes to compare the SoA arrays without a GPU.
bool g_commit = true;

struct Builder {
    Scene *scene;
    bool ok = true;
    explicit Builder(uint32_t reserve) : scene(new Scene) { scene->handle = r1_scene_create(reserve); }
    void camera(float fx, float fy, float fz, float vfov, float aperture, float focus, float ax = 0, float ay = 0, float az = 0)
    {
        const HostConfig &cfg = host_config();
        const float from[3] = { fx, fy, fz }, at[3] = { ax, ay, az }, up[3] = { 0, 1, 0 };
        // aspect = (float)SCREEN_W / (float)SCREEN_H (rayweek1.cpp:564)
        if (r1_scene_set_camera(scene->handle, from, at, up, vfov, (float)cfg.width / (float)cfg.height, aperture, focus)) ok = false;
    }
    void lambert(float x, float y, float z, float radius, float r, float g, float b) { add(x, y, z, radius, R1_MAT_LAMBERT, r, g, b, 0); }
    void metal(float x, float y, float z, float radius, float r, float g, float b, float fuzz) { add(x, y, z, radius, R1_MAT_METAL, r, g, b, fuzz); }
    void dielectric(float x, float y, float z, float radius, float ior) { add(x, y, z, radius, R1_MAT_DIELECTRIC, 1, 1, 1, ior); }
    void add(float x, float y, float z, float radius, int kind, float r, float g, float b, float param)
    {
        if (ok && r1_scene_add_sphere(scene->handle, x, y, z, radius, kind, r, g, b, param) < 0) ok = false;
    }
    Scene *finish()
    {
        // "make sure num spheres is multiple of SIMD width" (rayweek1.cpp:574-576); SIMD_WIDTH = 8 (:38)
        if (ok && r1_scene_pad(scene->handle, 8)) ok = false;
        // the builders emit the device buffers: one replica per GPU that will render
        const int n = host_config().n_gpus > 0 ? host_config().n_gpus : 1;
        for (int dev = n - 1; ok && g_commit && dev >= 0; --dev)
            if (r1_scene_commit(scene->handle, dev)) ok = false;
        if (!ok) { delete scene; return nullptr; }
        return scene;
    }
};

// rayweek1.cpp:668-712 for a gw x gh grid; ior_mod = 0 reproduces the reference (ior = 1.2 + 0.05 i)
Scene *grid_scene(int gw, int gh, int ior_mod, float fx, float fy, float fz, float focus)
{
    Builder b(gw * gh + 4 + 8);
    b.camera(fx, fy, fz, 60, 0.1f, focus);
    const int W = gw, H = gh;
    srand(111);
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            float px = (x - W / 2) * 1.1f, py = 0, pz = (y - H / 2) * 1.1f;
            // CRT random
            const float r = (rand() & 0xff) / 255.0f;
            const float g = (rand() & 0xff) / 255.0f;
            const float bl = (rand() & 0xff) / 255.0f;
            const int i = x + y * W;
            const float radius = 0.45f;
            if (i % 20 == 0) {
                const int k = ior_mod ? i % ior_mod : i;
                b.dielectric(px, py, pz, radius, 1.2f + k * 0.05f);
            } else if (i % 10 == 0) {
                py += 0.1f;
                b.metal(px, py, pz, radius, r, g, bl, 0.01f + 0.5f * y / (float)(H));
            } else {
                b.lambert(px, py, pz, radius, r, g, bl);
            }
        }
    }
    b.lambert(0, -1000.5f, 0, 1000, 0.5f, 0.5f, 0.5f);
    b.metal(5, 3, 0, 2, 0.5f, 0.5f, 0.8f, 0.65f);
    b.dielectric(0, 3, 0, 2, 1.5f);
    b.metal(-5, 3, 0, 2, 0.8f, 0.2f, 0.2f, 0.05f);
    return b.finish();
}

// rayweek1.cpp:552-579
Scene *build_small()
{
    Builder b(5 + 8);
    b.camera(2, 1, 2, 60, 0.1f, 5.0f);
    b.lambert(0, 0, -1, 0.5f, 0.1f, 0.2f, 0.5f);
    b.lambert(0, -100.5f, -1, 100.0f, 0.8f, 0.8f, 0);
    b.metal(1, 0, -1, 0.5f, 0.8f, 0.6f, 0.2f, 0.3f);
    b.dielectric(-1, 0, -1, 0.5f, 1.5f);
    b.dielectric(-1, 0, -1, -0.45f, 1.5f);
    return b.finish();
}

// rayweek1.cpp:582-651 ("the aras_p scene")
Scene *build_medium()
{
    Builder b(46 + 8);
    b.camera(0, 2, 3, 60, 0.1f * 0.2f, 3);
    b.lambert(0, -100.5, -1, 100, 0.8f, 0.8f, 0.8f);
    b.lambert(2, 0, -1, 0.5f, 0.8f, 0.4f, 0.4f);
    b.lambert(0, 0, -1, 0.5f, 0.4f, 0.8f, 0.4f);
    b.metal(-2, 0, -1, 0.5f, 0.4f, 0.4f, 0.8f, 0);
    b.metal(2, 0, 1, 0.5f, 0.4f, 0.8f, 0.4f, 0);
    b.metal(0, 0, 1, 0.5f, 0.4f, 0.8f, 0.4f, 0.2f);
    b.metal(-2, 0, 1, 0.5f, 0.4f, 0.8f, 0.4f, 0.6f);
    b.dielectric(0.5f, 1, 0.5f, 0.5f, 1.5f);
    b.lambert(-1.5f, 1.5f, 0.f, 0.3f, 0.8f, 0.6f, 0.2f);
    // rows z = -3 .. -6, x = 4 .. -4: lambert greys, metal greys, metal hues, lambert hues (the last sphere is metal)
    const float grey[9] = { 0.1f, 0.2f, 0.3f, 0.4f, 0.5f, 0.6f, 0.7f, 0.8f, 0.9f };
    const float hue[9][3] = { { 0.8f, 0.1f, 0.1f }, { 0.8f, 0.5f, 0.1f }, { 0.8f, 0.8f, 0.1f }, { 0.4f, 0.8f, 0.1f }, { 0.1f, 0.8f, 0.1f },
                              { 0.1f, 0.8f, 0.5f }, { 0.1f, 0.8f, 0.8f }, { 0.1f, 0.1f, 0.8f }, { 0.5f, 0.1f, 0.8f } };
    for (int k = 0; k < 9; ++k) b.lambert((float)(4 - k), 0, -3, 0.5f, grey[k], grey[k], grey[k]);
    for (int k = 0; k < 9; ++k) b.metal((float)(4 - k), 0, -4, 0.5f, grey[k], grey[k], grey[k], 0);
    for (int k = 0; k < 9; ++k) b.metal((float)(4 - k), 0, -5, 0.5f, hue[k][0], hue[k][1], hue[k][2], 0);
    for (int k = 0; k < 8; ++k) b.lambert((float)(4 - k), 0, -6, 0.5f, hue[k][0], hue[k][1], hue[k][2]);
    b.metal(-4, 0, -6, 0.5f, 0.5f, 0.1f, 0.8f, 0);
    b.lambert(1.5f, 1.5f, -2, 0.3f, 0.1f, 0.2f, 0.5f);
    return b.finish();
}

Scene *build_by_name(const char *name)
{
    if (!strcmp(name, "small")) return build_small();
    if (!strcmp(name, "medium")) return build_medium();
    if (!strcmp(name, "large")) return grid_scene(30, 16, 0, 3, 8, 15, 10.0f);          // rayweek1.cpp:654-719
    // 66 x 62 grid + 4 = 4096 spheres, ior kept inside the reference's range
    if (!strcmp(name, "synth4096")) return grid_scene(66, 62, 480, 6, 16, 30, 20.0f);
    return nullptr;
}

// Scene description files: what the reference hard-codes in create_*_scene(), as text, so that
// sphere-count sweeps do not need a recompile.  One statement per line, '#' starts a comment:
//   camera <from.xyz> <at.xyz> <vfov_deg> <aperture> <focus_dist>           (up is +y, aspect = width / height)
//   sphere <c.xyz> <radius> lambert <r> <g> <b> | metal <r> <g> <b> <fuzz> | dielectric <ior> | none
Scene *build_from_file(const char *path)
{
    FILE *f = fopen(path, "rt");
    if (!f) return nullptr;
    Builder b(64);
    bool have_camera = false;
    char line[512];
    int lineno = 0;
    while (b.ok && fgets(line, sizeof(line), f)) {
        ++lineno;
        if (char *hash = strchr(line, '#')) *hash = 0;
        char word[32], mat[32];
        float v[10];
        if (sscanf(line, "%31s", word) != 1) continue;
        if (!strcmp(word, "camera")) {
            if (sscanf(line, "%*s %f %f %f %f %f %f %f %f %f", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7], &v[8]) != 9) b.ok = false;
            else { b.camera(v[0], v[1], v[2], v[6], v[7], v[8], v[3], v[4], v[5]); have_camera = true; }
        } else if (!strcmp(word, "sphere")) {
            int n = 0;
            if (sscanf(line, "%*s %f %f %f %f %31s%n", &v[0], &v[1], &v[2], &v[3], mat, &n) != 5) { b.ok = false; break; }
            const char *rest = line + n;
            if (!strcmp(mat, "lambert") && sscanf(rest, "%f %f %f", &v[4], &v[5], &v[6]) == 3) b.lambert(v[0], v[1], v[2], v[3], v[4], v[5], v[6]);
            else if (!strcmp(mat, "metal") && sscanf(rest, "%f %f %f %f", &v[4], &v[5], &v[6], &v[7]) == 4) b.metal(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
            else if (!strcmp(mat, "dielectric") && sscanf(rest, "%f", &v[4]) == 1) b.dielectric(v[0], v[1], v[2], v[3], v[4]);
            else if (!strcmp(mat, "none")) b.add(v[0], v[1], v[2], v[3], R1_MAT_NONE, 0, 0, 0, 0);
            else b.ok = false;
        } else {
            b.ok = false;
        }
        if (!b.ok) fprintf(stderr, "rays1_b200: %s:%d: cannot parse '%s'\n", path, lineno, word);
    }
    fclose(f);
    if (!have_camera) b.ok = false;
    return b.finish();
}

Scene *must(Scene *s, const char *name)
{
    if (!s) die(name);
    return s;
}

}  // namespace

// The reference's builders (rayweek1.cpp:552, 582, 654): no error path there, fatal here.
Scene *create_small_scene() { return must(build_by_name("small"), "create_small_scene"); }
Scene *create_medium_scene() { return must(build_by_name("medium"), "create_medium_scene"); }
Scene *create_large_scene() { return must(build_by_name("large"), "create_large_scene"); }
Scene *create_synth4096_scene() { return must(build_by_name("synth4096"), "create_synth4096_scene"); }
Scene *create_scene_by_name(const char *name) { return build_by_name(name); }
Scene *create_scene_from_file(const char *path) { return build_from_file(path); }

// ------------------------------------------------------------------------------------------------ multi-GPU
// One process, G devices: every device renders its interleaved row tiles (r1_render_device, asynchronous), then
// exactly two NCCL collectives over NVLink: a framebuffer gather to device 0 (grouped ncclSend/ncclRecv) and an
// ncclReduce of the ray counters.  NCCL is dlopen'ed so that librays1_b200.so carries no link-time NCCL dependency
// (a Python process that already loaded torch's bundled libnccl.so.2 keeps using that one).
namespace {

struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
#define R1_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(lib, name)); if (!field) return false;
        R1_SYM(CommInitAll, "ncclCommInitAll") R1_SYM(CommDestroy, "ncclCommDestroy") R1_SYM(GroupStart, "ncclGroupStart")
        R1_SYM(GroupEnd, "ncclGroupEnd") R1_SYM(Send, "ncclSend") R1_SYM(Recv, "ncclRecv") R1_SYM(Reduce, "ncclReduce")
        R1_SYM(GetErrorString, "ncclGetErrorString")
#undef R1_SYM
        return true;
    }
};

struct MultiGpu {
    Nccl nccl;
    int n = 0;
    std::vector<ncclComm_t> comms;
    std::vector<cudaStream_t> streams;
    std::vector<uint8_t *> rgb;          // per device: local rows
    std::vector<unsigned long long *> rays, rays_sum;
    uint8_t *gathered = nullptr;         // device 0: n x stride
    uint8_t *final_img = nullptr;        // device 0: full image
    size_t stride = 0, final_bytes = 0;

    void check(cudaError_t e, const char *what)
    {
        if (e != cudaSuccess) { fprintf(stderr, "rays1_b200: %s: %s\n", what, cudaGetErrorString(e)); exit(1); }
    }
    void check(ncclResult_t r, const char *what)
    {
        if (r != ncclSuccess) { fprintf(stderr, "rays1_b200: %s: %s\n", what, nccl.GetErrorString(r)); exit(1); }
    }
    void init(int n_gpus)
    {
        if (n == n_gpus) return;
        if (n != 0) { fprintf(stderr, "rays1_b200: GPU count cannot change between renders (%d -> %d)\n", n, n_gpus); exit(1); }
        if (!nccl.load()) { fprintf(stderr, "rays1_b200: cannot load libnccl.so.2: %s\n", dlerror()); exit(1); }
        n = n_gpus;
        std::vector<int> devs(n);
        for (int i = 0; i < n; ++i) devs[i] = i;
        comms.resize(n);
        check(nccl.CommInitAll(comms.data(), n, devs.data()), "ncclCommInitAll");
        streams.resize(n); rgb.assign(n, nullptr); rays.assign(n, nullptr); rays_sum.assign(n, nullptr);
        for (int i = 0; i < n; ++i) {
            check(cudaSetDevice(i), "cudaSetDevice");
            check(cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking), "cudaStreamCreate");
            check(cudaMalloc(&rays[i], 8), "cudaMalloc");
            check(cudaMalloc(&rays_sum[i], 8), "cudaMalloc");
        }
    }
    void size_for(int width, int height, int row_tile)
    {
        size_t need = 0;
        for (int r = 0; r < n; ++r) need = std::max(need, (size_t)r1_local_pixels(width, height, row_tile, r, n) * 3);
        need = (need + 255) / 256 * 256;
        if (need > stride) {
            for (int i = 0; i < n; ++i) {
                check(cudaSetDevice(i), "cudaSetDevice");
                if (rgb[i]) cudaFree(rgb[i]);
                check(cudaMalloc(&rgb[i], need), "cudaMalloc");
            }
            check(cudaSetDevice(0), "cudaSetDevice");
            if (gathered) cudaFree(gathered);
            check(cudaMalloc(&gathered, need * n), "cudaMalloc");
            stride = need;
        }
        const size_t fb = (size_t)width * height * 3;
        if (fb > final_bytes) {
            check(cudaSetDevice(0), "cudaSetDevice");
            if (final_img) cudaFree(final_img);
            check(cudaMalloc(&final_img, fb), "cudaMalloc");
            final_bytes = fb;
        }
    }
};

MultiGpu &multi_gpu()
{
    static MultiGpu m;
    return m;
}

RESULT render_multi(Scene *scene, Pix *pixels, int n_gpus, double *kernel_ms)
{
    const HostConfig &cfg = host_config();
    MultiGpu &m = multi_gpu();
    m.init(n_gpus);
    m.size_for(cfg.width, cfg.height, cfg.row_tile);
    r1_render_params p;
    memset(&p, 0, sizeof(p));
    p.width = cfg.width; p.height = cfg.height; p.spp = cfg.spp; p.max_bounces = cfg.max_bounces;
    p.variant = cfg.variant; p.seed = cfg.seed; p.world = n_gpus; p.row_tile = cfg.row_tile;
    // 1. every device traces its row tiles (asynchronous launches from this one host thread)
    for (int r = 0; r < n_gpus; ++r) {
        p.rank = r; p.device = r;
        if (r1_render_device(scene->handle, &p, m.rgb[r], m.rays[r], m.streams[r], nullptr)) die("r1_render_device");
    }
    // 2. framebuffer gather to rank 0 + ray-counter reduce: the path's only exchange step
    m.check(m.nccl.GroupStart(), "ncclGroupStart");
    for (int r = 0; r < n_gpus; ++r) {
        const size_t bytes = (size_t)r1_local_pixels(cfg.width, cfg.height, cfg.row_tile, r, n_gpus) * 3;
        m.check(m.nccl.Send(m.rgb[r], bytes, ncclUint8, 0, m.comms[r], m.streams[r]), "ncclSend");
        m.check(m.nccl.Recv(m.gathered + (size_t)r * m.stride, bytes, ncclUint8, r, m.comms[0], m.streams[0]), "ncclRecv");
    }
    m.check(m.nccl.GroupEnd(), "ncclGroupEnd");
    m.check(m.nccl.GroupStart(), "ncclGroupStart");
    for (int r = 0; r < n_gpus; ++r)
        m.check(m.nccl.Reduce(m.rays[r], m.rays_sum[r], 1, ncclUint64, ncclSum, 0, m.comms[r], m.streams[r]), "ncclReduce");
    m.check(m.nccl.GroupEnd(), "ncclGroupEnd");
    // 3. de-interleave on device 0, one D2H of the RGB8 image
    m.check(cudaSetDevice(0), "cudaSetDevice");
    if (r1_deinterleave_rows(0, m.gathered, m.stride, m.final_img, cfg.width, cfg.height, cfg.row_tile, n_gpus, m.streams[0])) die("r1_deinterleave_rows");
    unsigned long long rays = 0;
    m.check(cudaMemcpyAsync(pixels, m.final_img, (size_t)cfg.width * cfg.height * 3, cudaMemcpyDeviceToHost, m.streams[0]), "cudaMemcpyAsync");
    m.check(cudaMemcpyAsync(&rays, m.rays_sum[0], 8, cudaMemcpyDeviceToHost, m.streams[0]), "cudaMemcpyAsync");
    for (int r = n_gpus - 1; r >= 0; --r) {
        m.check(cudaSetDevice(r), "cudaSetDevice");
        m.check(cudaStreamSynchronize(m.streams[r]), "cudaStreamSynchronize");
    }
    double worst = 0;
    for (int r = 0; r < n_gpus; ++r) {
        r1_result res;
        if (r1_render_wait(scene->handle, r, &res)) die("r1_render_wait");
        worst = std::max(worst, res.kernel_ms);
    }
    *kernel_ms = worst;
    RESULT out = { 0, rays, worst };
    return out;
}

}  // namespace

// rayweek1.cpp:845-927
RESULT benchmark(Scene *scene, Pix *pixels, bool write_tga, const char *scene_name)
{
    RESULT result = { 0, 0, 0 };
    const HostConfig &cfg = host_config();
    const auto t0 = std::chrono::steady_clock::now();  // Timer timer; (:848) -- scene construction is not timed

    const int n_gpus = cfg.n_gpus > 0 ? cfg.n_gpus : 1;
    if (n_gpus == 1) {
        r1_render_params p;
        memset(&p, 0, sizeof(p));
        p.width = cfg.width; p.height = cfg.height; p.spp = cfg.spp; p.max_bounces = cfg.max_bounces;
        p.variant = cfg.variant; p.seed = cfg.seed; p.rank = 0; p.world = 1; p.row_tile = cfg.row_tile; p.device = 0;
        r1_result res;
        if (r1_render(scene->handle, &p, reinterpret_cast<uint8_t *>(pixels), &res)) die("r1_render");
        result.num_rays = res.num_rays;
        result.kernel_ms = res.kernel_ms;
    } else {
        double kms = 0;
        result = render_multi(scene, pixels, n_gpus, &kms);
    }
    result.elapsed_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();  // :891

    const uint64_t total_samples = (uint64_t)cfg.width * (uint64_t)cfg.height * (uint64_t)cfg.spp;  // 64-bit, unlike :893
    if (!cfg.quiet) {
        int visible = r1_device_count();
        // the reference's block (:895-902), byte-compatible; "threads" = GPUs used / GPUs visible, "tile size" = row tile
        printf("%s\n", scene_name);
        printf("elapsed time:   %.3fs\n", result.elapsed_seconds);
        printf("total samples:  %llu\n", (unsigned long long)total_samples);
        printf("total rays:     %llu\n", (unsigned long long)result.num_rays);
        printf("mrays/s:        %0.2f\n", result.get_mrays_per_sec());
        printf("threads:        %d/%u\n", n_gpus, (unsigned)(visible > 0 ? visible : 0));
        printf("tile size:      %dx%d\n", cfg.width, cfg.row_tile);
        // extra lines go AFTER the reference's lines
        printf("gpu kernel ms:  %.3f\n", result.kernel_ms);
        printf("kernel mrays/s: %0.2f\n", result.kernel_ms > 0 ? result.num_rays / (result.kernel_ms * 1e-3) / 1e6 : 0.0);
        printf("\n");
    }

    delete scene;  // :905 -- benchmark() takes ownership

    if (write_tga) {  // :907-912
        char filename[128];
        snprintf(filename, sizeof(filename), "out_%s.tga", scene_name);
        tga_write_rgb24(filename, cfg.width, cfg.height, pixels);
    }
    return result;
}

// common.h:47-77
void log_results(const char *version, const char *scene, const RESULT *results, int num_runs)
{
    RESULT result = { 0, 0, 0 };
    for (int i = 0; i < num_runs; i++) {
        result.elapsed_seconds += results[i].elapsed_seconds;
        result.num_rays += results[i].num_rays;
    }
    result.elapsed_seconds /= num_runs;
    result.num_rays /= num_runs;
    char filename[128];
    snprintf(filename, sizeof(filename), "out_%s.txt", scene);
    FILE *f = fopen(filename, "wt");
    if (f) {
        fprintf(f, "%s|", version);
        fprintf(f, "%.3fs|", result.elapsed_seconds);
        fprintf(f, "%llu|", (long long unsigned)result.num_rays);
        fprintf(f, "%0.3f mrays/s|", result.get_mrays_per_sec());
        fclose(f);
    }
}

// common.h:86-122 -- 18-byte header, type 2, 24 bpp, descriptor 0 (bottom-left origin); swaps R and B in place
bool tga_write_rgb24(const char *filename, int width, int height, Pix *pixels)
{
    FILE *f = fopen(filename, "wb");
    if (!f) return false;
    uint8_t header[18] = { 0 };
    header[2] = 2;
    header[12] = (uint8_t)(width & 0x00FF);
    header[13] = (uint8_t)((width & 0xFF00) >> 8);
    header[14] = (uint8_t)(height & 0x00FF);
    header[15] = (uint8_t)((height & 0xFF00) >> 8);
    header[16] = 24;
    const size_t n = (size_t)width * height;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t tmp = pixels[i].r;
        pixels[i].r = pixels[i].b;
        pixels[i].b = tmp;
    }
    fwrite(header, 1, sizeof(header), f);
    fwrite(pixels, 3, n, f);
    fclose(f);
    return true;
}

// ------------------------------------------------------------------------------------------------ C linkage (part 2)
extern "C" {

int r1_host_configure(int width, int height, int spp, int max_bounces, int variant, int n_gpus, uint32_t seed)
{
    HostConfig &c = host_config();
    if (width > 0) c.width = width;
    if (height > 0) c.height = height;
    if (spp > 0) c.spp = spp;
    if (max_bounces > 0) c.max_bounces = max_bounces;
    if (variant >= 0) c.variant = variant;
    if (n_gpus > 0) c.n_gpus = n_gpus;
    c.seed = seed;
    return R1_OK;
}

int r1_host_set_quiet(int quiet)
{
    host_config().quiet = quiet != 0;
    return R1_OK;
}

void *r1_host_create_scene(const char *name, int commit)
{
    if (!name) return nullptr;
    g_commit = commit != 0;
    Scene *s = build_by_name(name);
    g_commit = true;
    return s;
}

void *r1_host_create_scene_from_file(const char *path, int commit)
{
    if (!path) return nullptr;
    g_commit = commit != 0;
    Scene *s = build_from_file(path);
    g_commit = true;
    return s;
}

r1_scene *r1_host_scene_handle(void *scene) { return scene ? static_cast<Scene *>(scene)->handle : nullptr; }

int r1_host_benchmark(void *scene, uint8_t *pixels, int write_tga, const char *scene_name, double *elapsed_seconds, uint64_t *num_rays,
                      double *kernel_ms)
{
    if (!scene || !pixels || !scene_name) return R1_ERR_ARG;
    const RESULT r = benchmark(static_cast<Scene *>(scene), reinterpret_cast<Pix *>(pixels), write_tga != 0, scene_name);
    if (elapsed_seconds) *elapsed_seconds = r.elapsed_seconds;
    if (num_rays) *num_rays = r.num_rays;
    if (kernel_ms) *kernel_ms = r.kernel_ms;
    return R1_OK;
}

void r1_host_destroy_scene(void *scene) { delete static_cast<Scene *>(scene); }

int r1_host_write_tga(const char *filename, int width, int height, uint8_t *pixels)
{
    if (!filename || !pixels || width <= 0 || height <= 0) return R1_ERR_ARG;
    return tga_write_rgb24(filename, width, height, reinterpret_cast<Pix *>(pixels)) ? R1_OK : R1_ERR_ARG;
}

int r1_host_log_results(const char *version, const char *scene, const double *elapsed_seconds, const uint64_t *num_rays, int num_runs)
{
    if (!version || !scene || !elapsed_seconds || !num_rays || num_runs <= 0 || num_runs > 1024) return R1_ERR_ARG;
    std::vector<RESULT> r(num_runs);
    for (int i = 0; i < num_runs; ++i) r[i] = RESULT{ elapsed_seconds[i], num_rays[i], 0 };
    log_results(version, scene, r.data(), num_runs);
    return R1_OK;
}

}  // extern "C"
