// ref_harness.cpp -- C entry points around the UNMODIFIED reference code (test infrastructure only).
//
// The reference sources are compiled from where they lie under /root/reference (never copied): this TU
// textually includes src/latest/rayweek1.cpp with `main` renamed, and oracle/Makefile links it with
// src/latest/soa_sphere.cpp into oracle/_ref/libref_rays1.so (git-ignored).  Every function below calls the
// reference's own Hitable::hit / Material::scatter / Camera::getRay / render_tile / TileRenderScheduler.
// Used (a) to pin oracle/rays1_oracle.c and to record tests/golden/*.npz (oracle/make_golden.py), and
// (b) as the CPU baseline ("kind": "reference") in bench.py.  Nothing in the product path links this.
//
// Build flags: -fno-access-control lets the harness read Dielectric::_refIdx (private, rayweek1.cpp:463);
// RTTI stays on so material kinds can be recovered with dynamic_cast.
#define main reference_main
#include "src/latest/rayweek1.cpp"
#undef main

#include <vector>

#define REF_API extern "C" __attribute__((visibility("default")))

struct ref_scene {
    Scene *scene;
};

#ifdef REF_SYNTH4096
// SURVEY.md 8d config 5, built with the reference's OWN classes (Scene, Hitable, SphereSOA::add, Camera::init, the three
// materials): a 66 x 62 grid + the ground sphere + three radius-2 spheres = 4096 spheres, the recipe of create_large_scene()
// (rayweek1.cpp:668-712) with ior = 1.2 + 0.05 (i % 480) so that it stays inside the reference's own range.  Only the build
// whose Hitable::hit holds 4096 spheres (oracle/Makefile: MAX_SPHERES patched in a temporary copy of rayweek1.cpp:174)
// can trace it.  The product's builder for the same scene is rays1_host.cpp:grid_scene(66, 62, 480, ...).
static Scene *harness_synth4096_scene()
{
    Scene *scene = new Scene;
    Hitable *world = new Hitable;
    scene->hitables = world;
    scene->camera.init(Vec3(6, 16, 30), Vec3(0, 0, 0), Vec3(0, 1, 0), 60, (float)SCREEN_W / (float)SCREEN_H, 0.1f, 20.0f);
    const int W = 66, H = 62;
    world->_soa_spheres.reserve(W * H + 4 + SIMD_WIDTH);
    srand(111);
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            Vec3 pos((x - W / 2) * 1.1f, 0, (y - H / 2) * 1.1f);
            const float r = (rand() & 0xff) / 255.0f;
            const float g = (rand() & 0xff) / 255.0f;
            const float b = (rand() & 0xff) / 255.0f;
            const int i = x + y * W;
            Material *m;
            if (i % 20 == 0) m = new Dielectric(1.2f + (i % 480) * 0.05f);
            else if (i % 10 == 0) { m = new Metal(Vec3(r, g, b), 0.01f + 0.5f * y / (float)(H)); pos += Vec3(0, 0.1f, 0); }
            else m = new Lambertian(Vec3(r, g, b));
            world->_soa_spheres.add(pos, 0.45f, m);
        }
    }
    world->_soa_spheres.add(Vec3(0, -1000.5f, 0), 1000, new Lambertian(Vec3(0.5f, 0.5f, 0.5f)));
    world->_soa_spheres.add(Vec3(5, 3, 0), 2, new Metal(Vec3(0.5f, 0.5f, 0.8f), 0.65f));
    world->_soa_spheres.add(Vec3(0, 3, 0), 2, new Dielectric(1.5f));
    world->_soa_spheres.add(Vec3(-5, 3, 0), 2, new Metal(Vec3(0.8f, 0.2f, 0.2f), 0.05f));
    while (world->_soa_spheres.getCount() % SIMD_WIDTH != 0) world->_soa_spheres.add(Vec3(999999999, 999999999, 999999999), 0, nullptr);
    return scene;
}
#endif

// largest sphere count Hitable::hit of THIS build can hold (rayweek1.cpp:174)
REF_API int ref_max_spheres(void)
{
#ifdef REF_SYNTH4096
    return 4096;
#else
    return 1024;
#endif
}

REF_API ref_scene *ref_scene_create(const char *name)
{
    Scene *s = nullptr;
    if (!strcmp(name, "small")) s = create_small_scene();
    else if (!strcmp(name, "medium")) s = create_medium_scene();
    else if (!strcmp(name, "large")) s = create_large_scene();
#ifdef REF_SYNTH4096
    else if (!strcmp(name, "synth4096")) s = harness_synth4096_scene();
#endif
    if (!s) return nullptr;
    return new ref_scene{ s };
}

REF_API void ref_scene_destroy(ref_scene *h)
{
    if (!h) return;
    delete h->scene;
    delete h;
}

REF_API uint32_t ref_scene_count(const ref_scene *h) { return h->scene->hitables->_soa_spheres.getCount(); }

static void material_flat(const Material *m, int32_t *kind, float *albedo, float *param)
{
    albedo[0] = albedo[1] = albedo[2] = 0;
    *param = 0;
    if (!m) { *kind = -1; return; }
    if (auto *l = dynamic_cast<const Lambertian *>(m)) {
        *kind = 0; albedo[0] = l->albedo.getX(); albedo[1] = l->albedo.getY(); albedo[2] = l->albedo.getZ();
    } else if (auto *me = dynamic_cast<const Metal *>(m)) {
        *kind = 1; albedo[0] = me->albedo.getX(); albedo[1] = me->albedo.getY(); albedo[2] = me->albedo.getZ();
        *param = me->fuzz;
    } else if (auto *d = dynamic_cast<const Dielectric *>(m)) {
        *kind = 2; albedo[0] = albedo[1] = albedo[2] = 1; *param = d->_refIdx;
    } else {
        *kind = -2;
    }
}

REF_API void ref_scene_get_soa(const ref_scene *h, float *cx, float *cy, float *cz, float *radius_sq, float *inv_radius,
                               int32_t *kind, float *albedo, float *param)
{
    const SphereSOA::InstanceData *d = h->scene->hitables->_soa_spheres.getData();
    for (uint32_t i = 0; i < d->_count; ++i) {
        cx[i] = d->center_x[i]; cy[i] = d->center_y[i]; cz[i] = d->center_z[i];
        radius_sq[i] = d->radius_sq[i]; inv_radius[i] = d->inv_radius[i];
        material_flat(d->material[i], &kind[i], &albedo[3 * i], &param[i]);
    }
}

static void put3(float *dst, const Vec3 &v) { dst[0] = v.getX(); dst[1] = v.getY(); dst[2] = v.getZ(); }

REF_API void ref_scene_get_camera(const ref_scene *h, float *out)
{
    const Camera &c = h->scene->camera;
    put3(out + 0, c._origin); put3(out + 3, c._lowerLeftCorner); put3(out + 6, c._horizontal); put3(out + 9, c._vertical);
    put3(out + 12, c._u); put3(out + 15, c._v); put3(out + 18, c._w);
    out[21] = c._lensRadius;
}

static int material_index(const ref_scene *h, const Material *m)
{
    const SphereSOA::InstanceData *d = h->scene->hitables->_soa_spheres.getData();
    for (uint32_t i = 0; i < d->_count; ++i)
        if (d->material[i] == m) return (int)i;
    return -1;
}

// Hitable::hit on a batch of rays whose directions are already unit length (set directly, no re-normalisation).
REF_API void ref_hit(const ref_scene *h, int n, const float *org, const float *dir, float t_min, float t_max,
                     int32_t *index, float *t, float *p, float *normal)
{
    for (int k = 0; k < n; ++k) {
        Ray r;
        r._origin = Vec3(org[3 * k], org[3 * k + 1], org[3 * k + 2]);
        r._dir = Vec3(dir[3 * k], dir[3 * k + 1], dir[3 * k + 2]);
        HitRecord rec;
        memset(&rec, 0, sizeof(rec));
        bool hit = h->scene->hitables->hit(r, t_min, t_max, &rec);
        index[k] = hit ? material_index(h, rec.material) : -1;
        t[k] = hit ? rec.t : 0.0f;
        if (hit) { put3(p + 3 * k, rec.p); put3(normal + 3 * k, rec.normal); }
        else { p[3 * k] = p[3 * k + 1] = p[3 * k + 2] = 0; normal[3 * k] = normal[3 * k + 1] = normal[3 * k + 2] = 0; }
    }
}

// Walk real paths with the reference's camera, hit() and scatter() and record every ray segment:
// ray (org, dir), depth, hit result, and for hits the scatter result together with the random inputs the
// reference's scatter consumed (obtained by replaying the generator on a COPY of the state just before the call).
// Returns the number of segments written (<= max_segments).
REF_API int ref_record_paths(const ref_scene *h, int max_segments, int image_w, int image_h, uint32_t seed,
                             float *org, float *dir, int32_t *depth_out, int32_t *index, float *t, float *p, float *normal,
                             float *rand_sphere, float *rand_u, int32_t *scat_ok, float *atten, float *scat_dir,
                             float *cam_su, float *cam_tv, float *cam_disk)
{
    ThreadData td;
    memset(&td, 0, sizeof(td));
    td.scene = h->scene;
    td.state = seed * 2u + 10001u;
    td.state4 = _mm_set_epi32(seed + 1001, seed + 1003, seed + 1005, seed + 1007);
    uint32_t pick = seed * 7919u + 12345u;
    int n = 0;
    while (n < max_segments) {
        int x = XorShift32(pick) % image_w, y = XorShift32(pick) % image_h;
        Vec3 uv = (Vec3(myrand01_x4(td.state4)) + Vec3((float)x, (float)y, 0)) * Vec3(1.0f / image_w, 1.0f / image_h, 0);
        uint32_t state_copy = td.state;
        Vec3 disk = random_in_unit_disk(state_copy);
        Ray r = h->scene->camera.getRay(uv.getX(), uv.getY(), td.state);
        for (int depth = 0; n < max_segments; ++depth) {
            put3(org + 3 * n, r._origin); put3(dir + 3 * n, r._dir);
            depth_out[n] = depth;
            cam_su[n] = uv.getX(); cam_tv[n] = uv.getY(); cam_disk[2 * n] = disk.getX(); cam_disk[2 * n + 1] = disk.getY();
            HitRecord rec;
            memset(&rec, 0, sizeof(rec));
            bool hit = h->scene->hitables->hit(r, 0.001f, FLT_MAX, &rec);
            index[n] = hit ? material_index(h, rec.material) : -1;
            t[n] = hit ? rec.t : 0.0f;
            for (int c = 0; c < 3; ++c) p[3 * n + c] = normal[3 * n + c] = rand_sphere[3 * n + c] = atten[3 * n + c] = scat_dir[3 * n + c] = 0;
            rand_u[n] = 0; scat_ok[n] = 0;
            if (!hit) { ++n; break; }
            put3(p + 3 * n, rec.p); put3(normal + 3 * n, rec.normal);
            if (depth >= MAX_BOUNCES) { ++n; break; } // color() does not call scatter past the cap (rayweek1.cpp:523)
            // replay the draws scatter() is about to make
            __m128i s4 = td.state4;
            put3(rand_sphere + 3 * n, random_in_unit_sphere(s4));
            uint32_t s1 = td.state;
            rand_u[n] = myrand01(s1);
            Vec3 attenuation(0, 0, 0);
            Ray scattered;
            scattered._origin = Vec3(0, 0, 0); scattered._dir = Vec3(0, 0, 0);
            bool ok = rec.material->scatter(r, rec, &attenuation, &scattered, &td);
            scat_ok[n] = ok ? 1 : 0;
            put3(atten + 3 * n, attenuation); put3(scat_dir + 3 * n, scattered._dir);
            ++n;
            if (!ok) break;
            r = scattered;
        }
    }
    return n;
}

// Hitable::hit followed by Material::scatter for GIVEN rays (one segment each, unit directions): the same record as
// ref_record_paths, for cases real camera paths reach rarely -- e.g. rays leaving the high-index dielectric spheres of the
// large scene from inside (ior up to 24.2, rayweek1.cpp:692).  The random inputs scatter() consumed are recovered by
// replaying the generators on copies of their states.
REF_API void ref_hit_scatter(const ref_scene *h, int n, const float *org, const float *dir, uint32_t seed, int32_t *index, float *t, float *p,
                             float *normal, float *rand_sphere, float *rand_u, int32_t *scat_ok, float *atten, float *scat_dir)
{
    ThreadData td;
    memset(&td, 0, sizeof(td));
    td.scene = h->scene;
    td.state = seed * 2u + 10001u;
    td.state4 = _mm_set_epi32(seed + 1001, seed + 1003, seed + 1005, seed + 1007);
    for (int k = 0; k < n; ++k) {
        Ray r;
        r._origin = Vec3(org[3 * k], org[3 * k + 1], org[3 * k + 2]);
        r._dir = Vec3(dir[3 * k], dir[3 * k + 1], dir[3 * k + 2]);
        HitRecord rec;
        memset(&rec, 0, sizeof(rec));
        const bool hit = h->scene->hitables->hit(r, 0.001f, FLT_MAX, &rec);
        index[k] = hit ? material_index(h, rec.material) : -1;
        t[k] = hit ? rec.t : 0.0f;
        for (int c = 0; c < 3; ++c) p[3 * k + c] = normal[3 * k + c] = rand_sphere[3 * k + c] = atten[3 * k + c] = scat_dir[3 * k + c] = 0;
        rand_u[k] = 0; scat_ok[k] = 0;
        if (!hit) continue;
        put3(p + 3 * k, rec.p); put3(normal + 3 * k, rec.normal);
        __m128i s4 = td.state4;
        put3(rand_sphere + 3 * k, random_in_unit_sphere(s4));
        uint32_t s1 = td.state;
        rand_u[k] = myrand01(s1);
        Vec3 attenuation(0, 0, 0);
        Ray scattered;
        scattered._origin = Vec3(0, 0, 0); scattered._dir = Vec3(0, 0, 0);
        const bool ok = rec.material->scatter(r, rec, &attenuation, &scattered, &td);
        scat_ok[k] = ok ? 1 : 0;
        put3(atten + 3 * k, attenuation); put3(scat_dir + 3 * k, scattered._dir);
    }
}

// Debug aid: ONE sample of ONE pixel from given generator states, walked with the reference's camera / hit / scatter and
// recorded segment by segment (same record as ref_record_paths).  Returns the number of segments (<= max_segments).
REF_API int ref_trace_sample(const ref_scene *h, int x, int y, int image_w, int image_h, uint32_t state, const uint32_t *state4, int max_segments,
                             float *org, float *dir, int32_t *index, float *t, int32_t *scat_ok)
{
    ThreadData td;
    memset(&td, 0, sizeof(td));
    td.scene = h->scene;
    td.state = state;
    td.state4 = _mm_loadu_si128((const __m128i *)state4);
    Vec3 uv = (Vec3(myrand01_x4(td.state4)) + Vec3((float)x, (float)y, 0)) * Vec3(1.0f / image_w, 1.0f / image_h, 0);
    Ray r = h->scene->camera.getRay(uv.getX(), uv.getY(), td.state);
    int n = 0;
    for (int depth = 0; n < max_segments; ++depth) {
        put3(org + 3 * n, r._origin); put3(dir + 3 * n, r._dir);
        HitRecord rec;
        memset(&rec, 0, sizeof(rec));
        bool hit = h->scene->hitables->hit(r, 0.001f, FLT_MAX, &rec);
        index[n] = hit ? material_index(h, rec.material) : -1;
        t[n] = hit ? rec.t : 0.0f;
        scat_ok[n] = 0;
        if (!hit || depth >= MAX_BOUNCES) { ++n; break; }
        Vec3 attenuation(0, 0, 0);
        Ray scattered;
        bool ok = rec.material->scatter(r, rec, &attenuation, &scattered, &td);
        scat_ok[n] = ok ? 1 : 0;
        ++n;
        if (!ok) break;
        r = scattered;
    }
    return n;
}

// Per-pixel replay fixtures: the pixel loop of render_tile (rayweek1.cpp:752-765) around the
// reference's OWN myrand01_x4 / Camera::getRay / color(), one pixel at a time.  For every pixel it records the two
// generator states BEFORE the pixel and the float radiance sum AFTER its spp samples, so that another implementation can
// replay each pixel independently from the same states (in the reference one flipped decision derails the shared
// sequential streams for every later pixel; recorded states confine it to one pixel).
REF_API void ref_replay_pixels(const ref_scene *h, int n, const int32_t *xy, int image_w, int image_h, int spp, uint32_t seed,
                               uint32_t *state_out, uint32_t *state4_out, float *color_sum, uint32_t *rays_out)
{
    ThreadData td;
    memset(&td, 0, sizeof(td));
    td.scene = h->scene;
    td.state = seed * 2u + 10001u;
    td.state4 = _mm_set_epi32(seed + 10001, seed + 10003, seed + 10005, seed + 10007);
    const Vec3 inv_image_size(1.0f / image_w, 1.0f / image_h, 0);
    for (int k = 0; k < n; ++k) {
        state_out[k] = td.state;
        _mm_storeu_si128((__m128i *)(state4_out + 4 * k), td.state4);
        const uint64_t rays_before = td.out_num_rays;
        Vec3 col(0, 0, 0);
        const Vec3 xyv((float)xy[2 * k], (float)xy[2 * k + 1], 0);
        for (int s = 0; s < spp; ++s) {
            Vec3 uv = (Vec3(myrand01_x4(td.state4)) + xyv) * inv_image_size;      // rayweek1.cpp:759
            Ray r = h->scene->camera.getRay(uv.getX(), uv.getY(), td.state);      // :760
            col += color(r, h->scene->hitables, 0, &td);                           // :762
        }
        put3(color_sum + 3 * k, col);
        rays_out[k] = (uint32_t)(td.out_num_rays - rays_before);
    }
}

// Camera::getRay with the disk sample recovered by replay (see ref_record_paths).
REF_API void ref_get_ray(const ref_scene *h, int n, const float *su, const float *tv, uint32_t seed, float *disk, float *org, float *dir)
{
    uint32_t state = seed;
    for (int k = 0; k < n; ++k) {
        uint32_t copy = state;
        Vec3 d = random_in_unit_disk(copy);
        disk[2 * k] = d.getX(); disk[2 * k + 1] = d.getY();
        Ray r = h->scene->camera.getRay(su[k], tv[k], state);
        put3(org + 3 * k, r._origin); put3(dir + 3 * k, r._dir);
    }
}

// benchmark() (rayweek1.cpp:845-891) with the image size / spp taken from arguments instead of the SCREEN_W/H macros:
// the reference's own TileRenderScheduler + render_tile do all the work.  threads <= 0 -> hardware_concurrency().
REF_API uint64_t ref_render(const ref_scene *h, uint8_t *rgb, int w, int h_px, int spp, int threads, double *elapsed_s)
{
    Timer timer;
    ThreadData td;
    memset(&td, 0, sizeof(td));
    td.scene = h->scene;
    td.image = (Pix *)rgb;
    td.image_w = w; td.image_h = h_px;
    td.tile_w_in_pixels = 32 < w ? 32 : w;
    td.tile_h_in_pixels = 32 < h_px ? 32 : h_px;
    td.samples_per_pixel = spp;
    td.out_num_rays = 0;
    int num_tiles = tiles_required(td.tile_w_in_pixels, w) * tiles_required(td.tile_h_in_pixels, h_px);
    int num_threads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    TileRenderScheduler scheduler;
    uint64_t rays = scheduler.run(num_tiles, num_threads, &td);
    if (elapsed_s) *elapsed_s = timer.elapsed();
    return rays;
}

REF_API int ref_hardware_concurrency(void) { return (int)std::thread::hardware_concurrency(); }

// RNG known answers
REF_API uint32_t ref_xorshift32(uint32_t *state) { return XorShift32(*state); }
REF_API float ref_myrand01(uint32_t *state) { return myrand01(*state); }
REF_API float ref_myrand02(uint32_t *state) { return myrand02(*state); }
REF_API void ref_myrand01_x4(uint32_t *state4, float *out)
{
    __m128i s = _mm_loadu_si128((const __m128i *)state4);
    _mm_storeu_ps(out, myrand01_x4(s));
    _mm_storeu_si128((__m128i *)state4, s);
}
REF_API void ref_random_in_unit_sphere(uint32_t *state4, float *out)
{
    __m128i s = _mm_loadu_si128((const __m128i *)state4);
    put3(out, random_in_unit_sphere(s));
    _mm_storeu_si128((__m128i *)state4, s);
}
REF_API void ref_random_in_unit_disk(uint32_t *state, float *out)
{
    Vec3 d = random_in_unit_disk(*state);
    out[0] = d.getX(); out[1] = d.getY();
}
