/* A non-Python host of the fused renderer, through the C-ABI only (include/perfb200.h + the CUDA runtime).
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/render_pano_host.c \
 *       -L perf_b200 -lperfb200 -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/perf_b200 -o render_pano_host
 *   ./render_pano_host geo_params.f32 app_params.f32 out.ppm        (flat fp32 tcnn params, 6 644 288 / 6 648 384 values:
 *                                                                    `scene.nerf.{geo,app}_mlp.params` of a PeRF checkpoint)
 *
 * Renders one 512 x 1024 panorama from the identity pose at 128 samples per ray (the inner loop of
 * CoreRunner.render_dense, core_exp_runner.py:229-238) and writes it as a binary PPM.  The same sequence of
 * calls is what a cgo / JNI / ctypes binding issues: cast the parameters once, pack the tables once, then one
 * perf_render_pano per frame.  tests/test_abi.py compiles and links this file; running it needs a B200. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include "perfb200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define CHECK_PERF(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s: %d (%s)\n", #x, rc_, perf_last_error()); return 1; } } while (0)

static float* read_floats(const char* path, uint64_t n)
{
    FILE* f = fopen(path, "rb");
    float* buf = (float*)malloc(n * sizeof(float));
    if (!f || !buf || fread(buf, sizeof(float), n, f) != n) { fprintf(stderr, "cannot read %llu floats from %s\n", (unsigned long long)n, path); exit(1); }
    fclose(f);
    return buf;
}

int main(int argc, char** argv)
{
    if (argc != 4) { fprintf(stderr, "usage: %s geo_params.f32 app_params.f32 out.ppm\n", argv[0]); return 2; }
    const perf_grid_cfg grid = {16, 2, 18, 16, 1.4472692012786865f, 0};      /* ngp_nerf.py:99-106 */
    const perf_mlp_cfg geo = {32, 1, 64, 1, 0}, app = {32, 3, 64, 2, 1};       /* ngp_nerf.py:107-113,127-133 */
    uint64_t n_geo = 0, n_app = 0, n_entries = 0;
    CHECK_PERF(perf_network_param_count(&grid, &geo, &n_geo));
    CHECK_PERF(perf_network_param_count(&grid, &app, &n_app));
    CHECK_PERF(perf_packed_table_entries(&grid, &n_entries));     /* entries of the packed gather table (grid + cell-major dense levels) */
    if (perf_device_arch() != 100) fprintf(stderr, "warning: built for sm_100a, device reports %d\n", perf_device_arch());

    const int H = 512, W = 1024, S = 128;
    float *h_geo = read_floats(argv[1], n_geo), *h_app = read_floats(argv[2], n_app);
    float *d_geo32, *d_app32, *d_rgb, *d_dist;
    void *d_geo16, *d_app16, *d_packed;
    cudaStream_t stream;
    CHECK_CUDA(cudaStreamCreate(&stream));
    CHECK_CUDA(cudaMalloc((void**)&d_geo32, n_geo * 4)); CHECK_CUDA(cudaMalloc((void**)&d_app32, n_app * 4));
    CHECK_CUDA(cudaMalloc(&d_geo16, n_geo * 2)); CHECK_CUDA(cudaMalloc(&d_app16, n_app * 2));
    CHECK_CUDA(cudaMalloc(&d_packed, n_entries * 8));
    CHECK_CUDA(cudaMalloc((void**)&d_rgb, (size_t)H * W * 3 * 4)); CHECK_CUDA(cudaMalloc((void**)&d_dist, (size_t)H * W * 4));
    CHECK_CUDA(cudaMemcpyAsync(d_geo32, h_geo, n_geo * 4, cudaMemcpyHostToDevice, stream));
    CHECK_CUDA(cudaMemcpyAsync(d_app32, h_app, n_app * 4, cudaMemcpyHostToDevice, stream));

    /* once per checkpoint: fp16 shadows, interleaved gather table */
    CHECK_PERF(perf_params_to_half(d_geo32, d_geo16, n_geo, stream));
    CHECK_PERF(perf_params_to_half(d_app32, d_app16, n_app, stream));
    CHECK_PERF(perf_pack_tables(&grid, &geo, &app, d_geo16, d_app16, d_packed, stream));

    /* once per frame */
    perf_render_args args;
    const float aabb[6] = {-1.f, -1.f, -1.f, 1.f, 1.f, 1.f};                   /* nerf.py:35 */
    const float pose[16] = {1, 0, 0, 0,  0, 1, 0, 0,  0, 0, 1, 0,  0, 0, 0, 1};
    args.grid = grid; args.d_packed_table = d_packed; args.d_geo_mlp_half = d_geo16; args.d_app_mlp_half = d_app16;
    for (int i = 0; i < 6; ++i) args.aabb[i] = aabb[i];
    args.n_samples = S; args.near = 1e-2f; args.far = 1.0f; args.flags = 0;
    args.d_jitter = NULL; args.d_bg_noise = NULL; args.d_rgb = d_rgb; args.d_distance = d_dist; args.d_opacity = NULL; args.image_width = 0;
    CHECK_PERF(perf_render_pano(&args, pose, H, W, 0, H, stream));

    float* h_rgb = (float*)malloc((size_t)H * W * 3 * 4);
    CHECK_CUDA(cudaMemcpyAsync(h_rgb, d_rgb, (size_t)H * W * 3 * 4, cudaMemcpyDeviceToHost, stream));
    CHECK_CUDA(cudaStreamSynchronize(stream));
    FILE* out = fopen(argv[3], "wb");
    if (!out) { perror(argv[3]); return 1; }
    fprintf(out, "P6\n%d %d\n255\n", W, H);
    for (size_t i = 0; i < (size_t)H * W * 3; ++i) {
        float v = h_rgb[i] < 0.f ? 0.f : (h_rgb[i] > 1.f ? 1.f : h_rgb[i]);
        fputc((int)(v * 255.f + 0.5f), out);
    }
    fclose(out);
    printf("wrote %s (%d x %d, %d samples per ray)\n", argv[3], W, H, S);
    return 0;
}
