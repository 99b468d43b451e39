/*
 * oracle/cpath.c -- plain-C (OpenMP) restatement of the eval-mode per-ray path, used as the
 * multi-threaded CPU baseline of bench.py and cross-checked against the PyTorch oracle.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): nothing under perf_b200/ links or calls this.
 *
 * Follows the same reference lines as the PyTorch oracle:
 *   ray position / sampling   modules/scene/nerf_renderer.py:127, oracle/sampler.py
 *   field                     modules/fields/ngp_nerf.py:136-162 (aabb normalise, selector, exp, sigmoid net)
 *   hash grid + MLP           tiny-cuda-nn 1.7 (not vendored): SURVEY.md Appendix A
 *   composite + background    modules/scene/nerf_renderer.py:170-197, nerfacc 0.5.3
 * Precision contract = oracle mixed mode: fp16 parameters, tcnn's fp16 fma blend, fp16 MLP operands with
 * fp32 accumulation, fp16 rounding of hidden activations and outputs, fp32 composite.
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC (oracle/cpath.py does it lazily, per host CPU).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef _Float16 half;
#define N_LEVELS 16
#define HID 64

typedef struct { float scale; uint32_t res, size, offset; int hashed; } level_t;

static inline float h2f(half h) { return (float)h; }
static inline half f2h(float f) { return (half)f; }
static inline float round_half(float f) { return (float)(half)f; }

/* tcnn GridEncodingTemplated constructor; scale pinned as in oracle/hashgrid.py::grid_scale */
static uint32_t build_levels(level_t* lv, int log2_hashmap_size, int base, float per_level_scale)
{
    const float log2s = (float)log2((double)per_level_scale);
    uint64_t offset = 0;
    for (int l = 0; l < N_LEVELS; ++l) {
        volatile float x = (float)l * log2s;
        const float scale = (float)(exp2((double)x) * (double)base - 1.0);
        const uint32_t res = (uint32_t)ceilf(scale) + 1u;
        uint64_t dense = (uint64_t)res * res * res;
        dense = (dense + 7) / 8 * 8;
        const uint64_t cap = 1ull << log2_hashmap_size;
        const uint64_t size = dense < cap ? dense : cap;
        uint64_t stride = 1; int dims = 0;
        while (dims < 3 && stride <= size) { stride *= res; ++dims; }
        lv[l].scale = scale; lv[l].res = res; lv[l].size = (uint32_t)size; lv[l].offset = (uint32_t)offset;
        lv[l].hashed = size < stride;
        offset += size;
    }
    return (uint32_t)offset;
}

static inline uint32_t grid_index(const level_t* L, uint32_t gx, uint32_t gy, uint32_t gz)
{
    uint32_t idx;
    if (L->hashed) idx = gx ^ (gy * 2654435761u) ^ (gz * 805459861u);
    else idx = gx + gy * L->res + gz * L->res * L->res;
    return idx % L->size;
}

/* 32 fp16 features of one point; table: fp16 [n_entries][2] */
static void encode(const level_t* lv, const half* table, float x, float y, float z, half* feat)
{
    for (int l = 0; l < N_LEVELS; ++l) {
        const level_t* L = &lv[l];
        const float px = fmaf(L->scale, x, 0.5f), py = fmaf(L->scale, y, 0.5f), pz = fmaf(L->scale, z, 0.5f);
        const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
        const uint32_t gx = (uint32_t)(int32_t)fx, gy = (uint32_t)(int32_t)fy, gz = (uint32_t)(int32_t)fz;
        const float wx = px - fx, wy = py - fy, wz = pz - fz;
        half a0 = 0, a1 = 0;
        for (int c = 0; c < 8; ++c) {
            const float w = (((c & 1) ? wx : 1.0f - wx) * ((c & 2) ? wy : 1.0f - wy)) * ((c & 4) ? wz : 1.0f - wz);
            const half* v = table + 2 * (size_t)(L->offset + grid_index(L, gx + (c & 1), gy + ((c >> 1) & 1), gz + ((c >> 2) & 1)));
            const float wh = h2f(f2h(w));
            a0 = f2h(wh * h2f(v[0]) + h2f(a0));       /* one fp16 rounding per fma (tcnn / HFMA2) */
            a1 = f2h(wh * h2f(v[1]) + h2f(a1));
        }
        feat[2 * l] = a0; feat[2 * l + 1] = a1;
    }
}

/* out = round_half(relu(W in)).  Wt: [K][HID] fp32 holding fp16-valued weights TRANSPOSED so the loop
 * over the 64 outputs vectorises without a reduction; `in` / `out`: fp16-valued floats; fp32 accumulate. */
static void layer(const float* restrict Wt, int K, const float* restrict in, float* restrict out)
{
    float acc[HID];
    for (int n = 0; n < HID; ++n) acc[n] = 0.f;
    for (int k = 0; k < K; ++k) {
        const float x = in[k];
        const float* w = Wt + (size_t)k * HID;
#pragma omp simd
        for (int n = 0; n < HID; ++n) acc[n] += w[n] * x;
    }
    for (int n = 0; n < HID; ++n) out[n] = round_half(acc[n] > 0.f ? acc[n] : 0.f);
}
static float out_dot(const float* restrict w, const float* restrict h)
{
    float acc = 0.f;
#pragma omp simd reduction(+:acc)
    for (int k = 0; k < HID; ++k) acc += w[k] * h[k];
    return round_half(acc);
}
/* [rows][K] fp16 row-major -> fp32 [K][rows] */
static float* transpose_to_float(const half* W, int rows, int K)
{
    float* t = (float*)malloc((size_t)rows * K * sizeof(float));
    if (t) for (int n = 0; n < rows; ++n) for (int k = 0; k < K; ++k) t[(size_t)k * rows + n] = h2f(W[(size_t)n * K + k]);
    return t;
}

/*
 * geo_params / app_params: fp32 flat tcnn params (6 644 288 / 6 648 384 values).  Eval-mode render of R
 * rays with S fixed samples in [near, far]; outputs rgb [R,3], distance [R], opacity [R].
 * Returns 0, or -1 on allocation failure.  n_threads <= 0: OpenMP default.
 */
int oracle_render_rays(const float* geo_params, const float* app_params, const float* rays_o, const float* rays_d,
                       long R, int S, float near, float far, float* rgb, float* dist, float* opacity, int n_threads)
{
    level_t lv[N_LEVELS];
    const uint32_t n_entries = build_levels(lv, 18, 16, 1.4472692012786865f);
    const size_t n_geo_mlp = 64 * 32 + 16 * 64, n_app_mlp = 64 * 32 + 64 * 64 + 16 * 64;
    const size_t n_geo = n_geo_mlp + 2 * (size_t)n_entries, n_app = n_app_mlp + 2 * (size_t)n_entries;
    half* g16 = (half*)malloc(n_geo * sizeof(half));
    half* a16 = (half*)malloc(n_app * sizeof(half));
    if (!g16 || !a16) { free(g16); free(a16); return -1; }
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n_geo; ++i) g16[i] = f2h(geo_params[i]);     /* params.to(half) */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n_app; ++i) a16[i] = f2h(app_params[i]);
    const half *gT = g16 + n_geo_mlp, *aT = a16 + n_app_mlp;
    float* gW1 = transpose_to_float(g16, HID, 32);
    float* aW1 = transpose_to_float(a16, HID, 32);
    float* aW2 = transpose_to_float(a16 + 64 * 32, HID, HID);
    float gWo[HID], aWo[3 * HID];
    for (int k = 0; k < HID; ++k) gWo[k] = h2f(g16[64 * 32 + k]);
    for (int k = 0; k < 3 * HID; ++k) aWo[k] = h2f(a16[64 * 32 + 64 * 64 + k]);
    if (!gW1 || !aW1 || !aW2) { free(g16); free(a16); free(gW1); free(aW1); free(aW2); return -1; }
    const float step = (far - near) / (float)S;
#pragma omp parallel for schedule(dynamic, 16)
    for (long r = 0; r < R; ++r) {
        const float* o = rays_o + 3 * r; const float* d = rays_d + 3 * r;
        float sum_sd = 0.f, W = 0.f, D = 0.f, C[3] = {0.f, 0.f, 0.f};
        for (int k = 0; k < S; ++k) {
            const float ts = near + (float)k * step, te = near + (float)(k + 1) * step;
            const float tsum = ts + te;
            float p[3], x01[3]; int inside = 1;
            for (int i = 0; i < 3; ++i) {
                p[i] = o[i] + (d[i] * tsum) * 0.5f;
                x01[i] = (p[i] + 1.0f) / 2.0f;                         /* aabb [-1,1]^3 (nerf.py:35) */
                inside &= (x01[i] > 0.f) & (x01[i] < 1.f);
            }
            float sigma = 0.f, c3[3] = {0.f, 0.f, 0.f};
            if (inside) {
                half feat[32]; float ff[32], h1[HID], h2[HID];
                encode(lv, gT, x01[0], x01[1], x01[2], feat);
                for (int i = 0; i < 32; ++i) ff[i] = h2f(feat[i]);
                layer(gW1, 32, ff, h1);
                sigma = expf(out_dot(gWo, h1));
                encode(lv, aT, x01[0], x01[1], x01[2], feat);
                for (int i = 0; i < 32; ++i) ff[i] = h2f(feat[i]);
                layer(aW1, 32, ff, h1);
                layer(aW2, HID, h1, h2);
                for (int i = 0; i < 3; ++i) c3[i] = round_half(1.0f / (1.0f + expf(-out_dot(aWo + i * HID, h2))));
            }
            const float sd = sigma * (te - ts);
            const float w = expf(-sum_sd) * (1.0f - expf(-sd));
            sum_sd += sd;
            W += w; D += w * (tsum * 0.5f);
            for (int i = 0; i < 3; ++i) C[i] += w * c3[i];
        }
        const float one_m = 1.0f - W;
        for (int i = 0; i < 3; ++i) rgb[3 * r + i] = C[i] + 0.5f * one_m;   /* nerf_renderer.py:195-197 */
        dist[r] = D + 5.0f * one_m;
        opacity[r] = W;
    }
    free(g16); free(a16); free(gW1); free(aW1); free(aW2);
    return 0;
}

int oracle_c_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
