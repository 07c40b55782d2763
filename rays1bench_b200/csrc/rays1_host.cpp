// rays1_host.cpp -- the reference's host surface (scene builders, benchmark(), report, TGA, log) on the C ABI.
// Mirrors src/latest/rayweek1.cpp:552-719, 845-927 and src/common/common.h:36-122 of the reference; the trace loop
// itself runs on the GPU(s) behind r1_render / r1_render_device.  file:line citations are relative to /root/reference/.
#include "rays1_host.h"
#include "r1_internal.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
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
            // CRT random.  The three constant expressions below (:679-681, :692, :696) are written the way the reference's
            // fast-math build (bench.py:175) evaluates them, so that albedo / ior / fuzz have the reference's exact bits:
            // x / 255.0f -> x * (1 / 255.0f);  1.2f + i * 0.05f -> one fma;  0.01f + 0.5f * y / H -> fma(0.5f * y, 1 / H, 0.01f)
            const float r = (rand() & 0xff) * (1.0f / 255.0f);
            const float g = (rand() & 0xff) * (1.0f / 255.0f);
            const float bl = (rand() & 0xff) * (1.0f / 255.0f);
            const int i = x + y * W;
            const float radius = 0.45f;
            if (i % 20 == 0) {
                const int k = ior_mod ? i % ior_mod : i;
                b.dielectric(px, py, pz, radius, fmaf((float)k, 0.05f, 1.2f));
            } else if (i % 10 == 0) {
                py += 0.1f;
                b.metal(px, py, pz, radius, r, g, bl, fmaf(0.5f * y, 1.0f / (float)(H), 0.01f));
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
// Everything here RETURNS an error code (message via r1_last_error()): the C-linkage entry points hand it to their
// caller, only the reference-shaped C++ benchmark() / create_*_scene() turn it into the fatal exit the reference's
// no-error-path surface implies.
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
    int load()
    {
        if (lib) return R1_OK;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return r1_set_error(R1_ERR_CUDA, "cannot load libnccl.so.2: %s", dlerror());
#define R1_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(h, name)); if (!field) { dlclose(h); return r1_set_error(R1_ERR_CUDA, "libnccl.so.2 lacks %s", name); }
        R1_SYM(CommInitAll, "ncclCommInitAll") R1_SYM(CommDestroy, "ncclCommDestroy") R1_SYM(GroupStart, "ncclGroupStart")
        R1_SYM(GroupEnd, "ncclGroupEnd") R1_SYM(Send, "ncclSend") R1_SYM(Recv, "ncclRecv") R1_SYM(Reduce, "ncclReduce")
        R1_SYM(GetErrorString, "ncclGetErrorString")
#undef R1_SYM
        lib = h;   // published only once every symbol resolved
        return R1_OK;
    }
};

#define R1H_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return r1_set_error(R1_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); } while (0)
#define R1H_NCCL(expr) do { ncclResult_t r_ = (expr); if (r_ != ncclSuccess) return r1_set_error(R1_ERR_CUDA, "%s: %s", #expr, nccl.GetErrorString(r_)); } while (0)
#define R1H_TRY(expr) do { int rc_ = (expr); if (rc_) return rc_; } while (0)

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

    void release()
    {
        for (int i = 0; i < n; ++i) {
            cudaSetDevice(i);
            if (i < (int)streams.size() && streams[i]) cudaStreamDestroy(streams[i]);
            if (i < (int)rgb.size() && rgb[i]) cudaFree(rgb[i]);
            if (i < (int)rays.size() && rays[i]) cudaFree(rays[i]);
            if (i < (int)rays_sum.size() && rays_sum[i]) cudaFree(rays_sum[i]);
            if (i < (int)comms.size() && comms[i]) nccl.CommDestroy(comms[i]);
        }
        if (n > 0) { cudaSetDevice(0); if (gathered) cudaFree(gathered); if (final_img) cudaFree(final_img); }
        comms.clear(); streams.clear(); rgb.clear(); rays.clear(); rays_sum.clear();
        gathered = final_img = nullptr; stride = final_bytes = 0; n = 0;
    }
    // communicator, streams, counters for n_gpus devices; a different GPU count than last time re-initialises
    int init(int n_gpus)
    {
        if (n == n_gpus) return R1_OK;
        release();
        R1H_TRY(nccl.load());
        int visible = 0;
        R1H_CUDA(cudaGetDeviceCount(&visible));
        if (n_gpus > visible) return r1_set_error(R1_ERR_ARG, "%d GPUs requested, %d visible", n_gpus, visible);
        std::vector<int> devs(n_gpus);
        for (int i = 0; i < n_gpus; ++i) devs[i] = i;
        comms.assign(n_gpus, nullptr);
        R1H_NCCL(nccl.CommInitAll(comms.data(), n_gpus, devs.data()));
        n = n_gpus;
        streams.assign(n, nullptr); rgb.assign(n, nullptr); rays.assign(n, nullptr); rays_sum.assign(n, nullptr);
        for (int i = 0; i < n; ++i) {
            R1H_CUDA(cudaSetDevice(i));
            R1H_CUDA(cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking));
            R1H_CUDA(cudaMalloc(&rays[i], 8));
            R1H_CUDA(cudaMalloc(&rays_sum[i], 8));
        }
        return R1_OK;
    }
    int size_for(int width, int height, int row_tile)
    {
        size_t need = 0;
        for (int r = 0; r < n; ++r) need = std::max(need, (size_t)r1_local_pixels(width, height, row_tile, r, n) * 3);
        need = (need + 255) / 256 * 256;
        if (need > stride) {
            for (int i = 0; i < n; ++i) {
                R1H_CUDA(cudaSetDevice(i));
                if (rgb[i]) cudaFree(rgb[i]);
                rgb[i] = nullptr;
                R1H_CUDA(cudaMalloc(&rgb[i], need));
            }
            R1H_CUDA(cudaSetDevice(0));
            if (gathered) cudaFree(gathered);
            gathered = nullptr; stride = 0;
            R1H_CUDA(cudaMalloc(&gathered, need * n));
            stride = need;
        }
        const size_t fb = (size_t)width * height * 3;
        if (fb > final_bytes) {
            R1H_CUDA(cudaSetDevice(0));
            if (final_img) cudaFree(final_img);
            final_img = nullptr; final_bytes = 0;
            R1H_CUDA(cudaMalloc(&final_img, fb));
            final_bytes = fb;
        }
        return R1_OK;
    }
    int render(Scene *scene, Pix *pixels, uint64_t *num_rays, double *kernel_ms);
};

MultiGpu &multi_gpu()
{
    static MultiGpu m;
    return m;
}

// replaces TileRenderScheduler::run (rayweek1.cpp:788-842) for n > 1: trace on every device, gather, reduce, one D2H
int MultiGpu::render(Scene *scene, Pix *pixels, uint64_t *num_rays, double *kernel_ms)
{
    const HostConfig &cfg = host_config();
    r1_render_params p;
    memset(&p, 0, sizeof(p));
    p.width = cfg.width; p.height = cfg.height; p.spp = cfg.spp; p.max_bounces = cfg.max_bounces;
    p.variant = cfg.variant; p.seed = cfg.seed; p.world = n; p.row_tile = cfg.row_tile;
    // 1. every device traces its row tiles (asynchronous launches from this one host thread)
    for (int r = 0; r < n; ++r) {
        p.rank = r; p.device = r;
        R1H_TRY(r1_render_device(scene->handle, &p, rgb[r], rays[r], streams[r], nullptr));
    }
    // 2. framebuffer gather to rank 0 + ray-counter reduce: the path's only exchange step
    R1H_NCCL(nccl.GroupStart());
    for (int r = 0; r < n; ++r) {
        const size_t bytes = (size_t)r1_local_pixels(cfg.width, cfg.height, cfg.row_tile, r, n) * 3;
        R1H_NCCL(nccl.Send(rgb[r], bytes, ncclUint8, 0, comms[r], streams[r]));
        R1H_NCCL(nccl.Recv(gathered + (size_t)r * stride, bytes, ncclUint8, r, comms[0], streams[0]));
    }
    R1H_NCCL(nccl.GroupEnd());
    R1H_NCCL(nccl.GroupStart());
    for (int r = 0; r < n; ++r) R1H_NCCL(nccl.Reduce(rays[r], rays_sum[r], 1, ncclUint64, ncclSum, 0, comms[r], streams[r]));
    R1H_NCCL(nccl.GroupEnd());
    // 3. de-interleave on device 0, one D2H of the RGB8 image
    R1H_CUDA(cudaSetDevice(0));
    R1H_TRY(r1_deinterleave_rows(0, gathered, stride, final_img, cfg.width, cfg.height, cfg.row_tile, n, streams[0]));
    unsigned long long total = 0;
    R1H_CUDA(cudaMemcpyAsync(pixels, final_img, (size_t)cfg.width * cfg.height * 3, cudaMemcpyDeviceToHost, streams[0]));
    R1H_CUDA(cudaMemcpyAsync(&total, rays_sum[0], 8, cudaMemcpyDeviceToHost, streams[0]));
    for (int r = n - 1; r >= 0; --r) {
        R1H_CUDA(cudaSetDevice(r));
        R1H_CUDA(cudaStreamSynchronize(streams[r]));
    }
    double worst = 0;
    for (int r = 0; r < n; ++r) {
        r1_result res;
        R1H_TRY(r1_render_wait(scene->handle, r, &res));
        worst = std::max(worst, res.kernel_ms);
    }
    *kernel_ms = worst;
    *num_rays = total;
    return R1_OK;
}

double g_last_kernel_ms = 0;

// benchmark() with an error path.  Takes ownership of the scene in every case (rayweek1.cpp:905).
int benchmark_impl(Scene *scene, Pix *pixels, size_t pixels_bytes, bool write_tga, const char *scene_name, RESULT *out)
{
    struct Owner { Scene *s; ~Owner() { delete s; } } owner{ scene };
    const HostConfig &cfg = host_config();
    const size_t need = (size_t)cfg.width * (size_t)cfg.height * sizeof(Pix);
    if (pixels_bytes < need)
        return r1_set_error(R1_ERR_ARG, "pixels holds %zu bytes, the configured %dx%d image needs %zu", pixels_bytes, cfg.width, cfg.height, need);
    const int n_gpus = cfg.n_gpus > 0 ? cfg.n_gpus : 1;
    // Set-up that the reference does not have (NCCL communicator, per-device staging) happens BEFORE the timer: the reference
    // times worker spawn + render + join (rayweek1.cpp:848-891), never one-off initialisation.
    if (n_gpus > 1) {
        R1H_TRY(multi_gpu().init(n_gpus));
        R1H_TRY(multi_gpu().size_for(cfg.width, cfg.height, cfg.row_tile));
    }
    RESULT result = { 0, 0 };
    double kernel_ms = 0;
    const auto t0 = std::chrono::steady_clock::now();  // Timer timer; (:848) -- scene construction is not timed
    if (n_gpus == 1) {
        r1_render_params p;
        memset(&p, 0, sizeof(p));
        p.width = cfg.width; p.height = cfg.height; p.spp = cfg.spp; p.max_bounces = cfg.max_bounces;
        p.variant = cfg.variant; p.seed = cfg.seed; p.rank = 0; p.world = 1; p.row_tile = cfg.row_tile; p.device = 0;
        r1_result res;
        R1H_TRY(r1_render(scene->handle, &p, reinterpret_cast<uint8_t *>(pixels), &res));
        result.num_rays = res.num_rays;
        kernel_ms = res.kernel_ms;
    } else {
        R1H_TRY(multi_gpu().render(scene, pixels, &result.num_rays, &kernel_ms));
    }
    result.elapsed_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();  // :891
    g_last_kernel_ms = kernel_ms;

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
        printf("gpu kernel ms:  %.3f\n", kernel_ms);
        printf("kernel mrays/s: %0.2f\n", kernel_ms > 0 ? result.num_rays / (kernel_ms * 1e-3) / 1e6 : 0.0);
        printf("\n");
    }
    if (write_tga) {  // :907-912 (the scene is deleted first there too, :905)
        char filename[128];
        snprintf(filename, sizeof(filename), "out_%s.tga", scene_name);
        tga_write_rgb24(filename, cfg.width, cfg.height, pixels);
    }
    *out = result;
    return R1_OK;
}

}  // namespace

// rayweek1.cpp:845-927
RESULT benchmark(Scene *scene, Pix *pixels, bool write_tga, const char *scene_name)
{
    RESULT result = { 0, 0 };
    const HostConfig &cfg = host_config();
    // the reference's caller owns SCREEN_W * SCREEN_H pixels (:960); there is no size to check on this surface
    if (benchmark_impl(scene, pixels, (size_t)cfg.width * cfg.height * sizeof(Pix), write_tga, scene_name, &result)) die("benchmark");
    return result;
}

double benchmark_last_kernel_ms() { return g_last_kernel_ms; }

// common.h:47-77
void log_results(const char *version, const char *scene, const RESULT *results, int num_runs)
{
    RESULT result = { 0, 0 };
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

int r1_host_benchmark(void *scene, uint8_t *pixels, uint64_t pixels_bytes, int write_tga, const char *scene_name, double *elapsed_seconds,
                      uint64_t *num_rays, double *kernel_ms)
{
    if (!scene) return r1_set_error(R1_ERR_ARG, "null scene");
    if (!pixels || !scene_name) { delete static_cast<Scene *>(scene); return r1_set_error(R1_ERR_ARG, "null argument"); }
    RESULT r = { 0, 0 };
    const int rc = benchmark_impl(static_cast<Scene *>(scene), reinterpret_cast<Pix *>(pixels), (size_t)pixels_bytes, write_tga != 0, scene_name, &r);
    if (rc) return rc;
    if (elapsed_seconds) *elapsed_seconds = r.elapsed_seconds;
    if (num_rays) *num_rays = r.num_rays;
    if (kernel_ms) *kernel_ms = g_last_kernel_ms;
    return R1_OK;
}

void r1_host_destroy_scene(void *scene) { delete static_cast<Scene *>(scene); }

int r1_host_write_tga(const char *filename, int width, int height, uint8_t *pixels, uint64_t pixels_bytes)
{
    if (!filename || !pixels || width <= 0 || height <= 0 || width > 0xFFFF || height > 0xFFFF) return r1_set_error(R1_ERR_ARG, "bad argument");
    if (pixels_bytes < (uint64_t)width * (uint64_t)height * 3) return r1_set_error(R1_ERR_ARG, "pixels holds %llu bytes, %dx%d needs %llu",
                                                                                   (unsigned long long)pixels_bytes, width, height, (unsigned long long)width * height * 3);
    return tga_write_rgb24(filename, width, height, reinterpret_cast<Pix *>(pixels)) ? R1_OK : r1_set_error(R1_ERR_ARG, "cannot write %s", filename);
}

int r1_host_log_results(const char *version, const char *scene, const double *elapsed_seconds, const uint64_t *num_rays, int num_runs)
{
    if (!version || !scene || !elapsed_seconds || !num_rays || num_runs <= 0 || num_runs > 1024) return R1_ERR_ARG;
    std::vector<RESULT> r(num_runs);
    for (int i = 0; i < num_runs; ++i) r[i] = RESULT{ elapsed_seconds[i], num_rays[i] };
    log_results(version, scene, r.data(), num_runs);
    return R1_OK;
}

}  // extern "C"
