// occ.cu -- occupancy-grid interval sampler (nerfacc OccGridEstimator.sampling semantics for
// levels=1, cone_angle=0; call site modules/scene/nerf_renderer.py:145-155).
// Rule (restated in oracle/occ_sampler.py): lattice t_k = near + (k + u_r) * step; the interval
// [t_k, t_k + step) is emitted when its midpoint lies inside the ray/aabb overlap and in an occupied
// cell.  Two passes (count, write) around an exclusive scan of the per-ray counts done by the caller.
#include "common.cuh"

namespace perf {

constexpr uint32_t PERF_OCC_MAX_STEPS = 1u << 22;      // lattice points one ray may visit (PeRF: 1.5 / 5e-4 = 3000)

struct OccArgs {
    const uint8_t* binaries; int rx, ry, rz;
    float amin[3], aext[3], amax[3];
    const float *rays_o, *rays_d, *jitter;
    uint64_t R; float near, far, step;
    int32_t* counts; const int64_t* offsets;
    int64_t* ray_indices; float *t_starts, *t_ends;
};

// One ray.  __host__ __device__: tests/host_harness.py compiles this file with -DPERF_HOST_HARNESS into a separate
// test-only object and runs the same body over host arrays against oracle/occ_sampler.py.
template <bool WRITE>
__host__ __device__ __forceinline__ void occ_march_ray(const OccArgs& a, uint64_t ray)
{
    const float o[3] = {a.rays_o[3 * ray], a.rays_o[3 * ray + 1], a.rays_o[3 * ray + 2]};
    const float d[3] = {a.rays_d[3 * ray], a.rays_d[3 * ray + 1], a.rays_d[3 * ray + 2]};
    float tn = -INFINITY, tf = INFINITY;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float inv = PERF_FDIV_RN(1.0f, fabsf(d[i]) < 1e-12f ? 1e-12f : d[i]);
        const float t0 = PERF_FMUL_RN(PERF_FSUB_RN(a.amin[i], o[i]), inv), t1 = PERF_FMUL_RN(PERF_FSUB_RN(a.amax[i], o[i]), inv);
        tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    }
    tn = fmaxf(tn, a.near); tf = fminf(tf, a.far);
    const float u = a.jitter ? a.jitter[ray] : 0.f;
    const float half_step = PERF_FMUL_RN(0.5f, a.step);
    int64_t pos = WRITE ? a.offsets[ray] : 0;
    int32_t n = 0;
    if (tf >= tn) {
        // first lattice index whose midpoint can reach tn (minus a safety margin; exact test below)
        float kf = floorf((tn - a.near) / a.step - u - 0.5f) - 2.0f;
        uint32_t k = kf > 0.f ? (uint32_t)kf : 0u;
        // Bounded walk: at most the lattice points between tn and tf (+ margin).  A ray whose overlap with the box
        // is not bounded by the box (|d| ~ 0 with far_plane = 1e10, or a tiny step) would otherwise spin once
        // (float)k stops changing at 2^24; such a ray yields no samples.
        const float span = (tf - tn) / a.step;
        const uint32_t k_end = span < (float)PERF_OCC_MAX_STEPS ? k + (uint32_t)span + 8u : k;
        for (; k < k_end; ++k) {
            const float ts = PERF_FADD_RN(a.near, PERF_FMUL_RN(PERF_FADD_RN((float)k, u), a.step));
            const float mid = PERF_FADD_RN(ts, half_step);
            if (mid > tf) break;
            if (mid < tn) continue;
            int c[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float p = PERF_FADD_RN(o[i], PERF_FMUL_RN(d[i], mid));
                const int res = i == 0 ? a.rx : (i == 1 ? a.ry : a.rz);
                int ci = (int)floorf(PERF_FMUL_RN(PERF_FDIV_RN(PERF_FSUB_RN(p, a.amin[i]), a.aext[i]), (float)res));
                c[i] = ci < 0 ? 0 : (ci > res - 1 ? res - 1 : ci);
            }
            if (a.binaries[((int64_t)c[0] * a.ry + c[1]) * a.rz + c[2]]) {
                if (WRITE) { a.ray_indices[pos] = (int64_t)ray; a.t_starts[pos] = ts; a.t_ends[pos] = PERF_FADD_RN(ts, a.step); ++pos; }
                ++n;
            }
        }
    }
    if (!WRITE) a.counts[ray] = n;
}

template <bool WRITE>
__global__ void __launch_bounds__(128) occ_march_kernel(const OccArgs a)
{
    const uint64_t ray = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray < a.R) occ_march_ray<WRITE>(a, ray);
}

}  // namespace perf

using namespace perf;

static int fill(OccArgs& a, const uint8_t* bin, const int* res3, const float* aabb6, const float* o, const float* d, const float* jit,
                uint64_t R, float near, float far, float step)
{
    PERF_CHECK_ARG(bin && res3 && aabb6 && o && d, "NULL pointer");
    PERF_CHECK_ARG(res3[0] > 0 && res3[1] > 0 && res3[2] > 0 && step > 0.f && far > near, "bad occupancy sampling arguments");
    memset(&a, 0, sizeof(a));
    a.binaries = bin; a.rx = res3[0]; a.ry = res3[1]; a.rz = res3[2];
    for (int i = 0; i < 3; ++i) { a.amin[i] = aabb6[i]; a.amax[i] = aabb6[3 + i]; a.aext[i] = aabb6[3 + i] - aabb6[i]; }
    a.rays_o = o; a.rays_d = d; a.jitter = jit; a.R = R; a.near = near; a.far = far; a.step = step;
    return PERF_OK;
}

extern "C" {
#pragma GCC visibility push(default)

int perf_occ_count(const uint8_t* d_binaries, const int* h_res3, const float* h_aabb6, const float* d_rays_o, const float* d_rays_d,
                   const float* d_jitter, uint64_t R, float near, float far, float step, int32_t* d_counts, void* stream)
{
    OccArgs a; int rc = fill(a, d_binaries, h_res3, h_aabb6, d_rays_o, d_rays_d, d_jitter, R, near, far, step); if (rc) return rc;
    PERF_CHECK_ARG(d_counts, "NULL counts");
    a.counts = d_counts;
    if (R == 0) return PERF_OK;
    occ_march_kernel<false><<<(unsigned)((R + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    PERF_LAUNCH_CHECK();
    return PERF_OK;
}

int perf_occ_write(const uint8_t* d_binaries, const int* h_res3, const float* h_aabb6, const float* d_rays_o, const float* d_rays_d,
                   const float* d_jitter, uint64_t R, float near, float far, float step, const int64_t* d_offsets,
                   int64_t* d_ray_indices, float* d_t_starts, float* d_t_ends, void* stream)
{
    OccArgs a; int rc = fill(a, d_binaries, h_res3, h_aabb6, d_rays_o, d_rays_d, d_jitter, R, near, far, step); if (rc) return rc;
    PERF_CHECK_ARG(d_offsets && d_ray_indices && d_t_starts && d_t_ends, "NULL output");
    a.offsets = d_offsets; a.ray_indices = d_ray_indices; a.t_starts = d_t_starts; a.t_ends = d_t_ends;
    if (R == 0) return PERF_OK;
    occ_march_kernel<true><<<(unsigned)((R + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    PERF_LAUNCH_CHECK();
    return PERF_OK;
}

#ifdef PERF_HOST_HARNESS
/* TEST HARNESS ONLY (never compiled into libperfb200.so): the per-ray body over HOST arrays.  pass 0 = count, 1 = write. */
int perf_host_occ_march(int pass, const uint8_t* h_binaries, const int* h_res3, const float* h_aabb6, const float* h_rays_o,
                        const float* h_rays_d, const float* h_jitter, uint64_t R, float near, float far, float step,
                        int32_t* h_counts, const int64_t* h_offsets, int64_t* h_ray_indices, float* h_t_starts, float* h_t_ends)
{
    OccArgs a; int rc = fill(a, h_binaries, h_res3, h_aabb6, h_rays_o, h_rays_d, h_jitter, R, near, far, step); if (rc) return rc;
    a.counts = h_counts; a.offsets = h_offsets; a.ray_indices = h_ray_indices; a.t_starts = h_t_starts; a.t_ends = h_t_ends;
    for (uint64_t r = 0; r < R; ++r) { if (pass == 0) occ_march_ray<false>(a, r); else occ_march_ray<true>(a, r); }
    return PERF_OK;
}
#endif

#pragma GCC visibility pop
}
