// packed.cu -- the occupancy-sampler (packed, variable samples per ray) half of the fused training step and the
// occupancy-grid update (VERDICT r1 missing #2 / #3).
//
// PeRF trains with `estimator_type: occ` (configs/nerf.yaml:25): nerfacc's sampler emits packed intervals sorted by
// ray, evaluates the density once under no_grad to drop samples behind transmittance 1e-4, and the renderer then
// evaluates both networks on the survivors (nerf_renderer.py:145-183).  Here one step is
//     perf_occ_count / perf_occ_write      (csrc/occ.cu)      intervals, all of them
//     perf_fields_packed                   (csrc/render.cu)   both fields at every interval + saves, ONE evaluation
//     perf_composite_packed_fwd            (this file)        transmittance scan with the 1e-4 cut applied inside
//     perf_composite_packed_bwd            (this file)        dL/d(pre-activation) of the trained network
//     perf_mlp_bwd                         (csrc/mlp_bwd.cu)
//     perf_hashgrid_bwd_merged             (csrc/api_basic.cu) same-cell runs merged before the atomics
// Dropping a sample whose transmittance is below the threshold and giving it weight zero are the same thing --
// transmittance is non-increasing along the ray, so the dropped samples are a suffix and do not enter anyone's prefix
// sum -- which is what lets the extra density evaluation and the stream compaction of the reference go away.
// Arithmetic contract: oracle/composite.py (render_weight_from_density, accumulate_along_rays, flatten_eff_distloss).
#include "common.cuh"

namespace perf {

__device__ __forceinline__ float wscan(float v, int lane)          // inclusive warp scan
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const float t = __shfl_up_sync(0xffffffffu, v, off); if (lane >= off) v += t; }
    return v;
}
__device__ __forceinline__ float wsum(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ float wlast(float v) { return __shfl_sync(0xffffffffu, v, 31); }

struct PkCompArgs {
    const int64_t* offsets;      // [R+1]
    const float *ts, *te, *sigma; const __half* rgb;      // [N], [N], [N], [N,4]
    uint64_t R; float eps; int training;
    const float* bg_noise;       // [R,4] or null
    float *w, *T;                // [N] saves (T = 0 marks a dropped sample)
    float *rgb_out, *dist_out, *op_out, *dacc, *dl;       // [R,3], [R], [R], [R], [R]
    // backward
    const float *g_rgb, *g_dist, *g_op, *g_dl;
    float* dz;                   // [N] (density phase) or [N,3] (colour phase)
};

// one warp per ray, lanes over 32 consecutive samples, chunks in order
__global__ void __launch_bounds__(256) composite_packed_fwd_kernel(const PkCompArgs a)
{
    const int lane = threadIdx.x & 31;
    const uint64_t ray = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= a.R) return;
    const int64_t b0 = a.offsets[ray], b1 = a.offsets[ray + 1];
    float carry_sd = 0.f, W = 0.f, D = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, du = 0.f, db = 0.f;
    for (int64_t base = b0; base < b1; base += 32) {
        const int64_t n = base + lane;
        const bool live = n < b1;
        float ts = 0.f, te = 0.f, sig = 0.f, r = 0.f, g = 0.f, b = 0.f;
        if (live) {
            ts = a.ts[n]; te = a.te[n]; sig = a.sigma[n];
            const uint2 c = *reinterpret_cast<const uint2*>(a.rgb + n * 4);
            const float2 c01 = unpack_half2(c.x), c2 = unpack_half2(c.y);
            r = c01.x; g = c01.y; b = c2.x;
        }
        const float dt = __fsub_rn(te, ts), m = __fadd_rn(ts, te) * 0.5f;
        const float sd = sig * dt;
        const float incl = wscan(sd, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1); if (lane == 0) excl = 0.f;
        const float T = expf(-(carry_sd + excl));
        const bool alive = live && T >= a.eps;                         // nerfacc early_stop_eps: later samples are dropped
        const float w = alive ? T * (1.f - expf(-sd)) : 0.f;
        if (live) { a.w[n] = w; a.T[n] = alive ? T : 0.f; }
        const float wm = w * m;
        const float wi = wscan(w, lane), wmi = wscan(wm, lane);
        const float Wx = W + (wi - w), WMx = D + (wmi - wm);           // exclusive prefix sums along the ray
        du += wsum(dt * w * w);
        db += wsum(w * (m * Wx - WMx));
        W += wlast(wi); D += wlast(wmi);
        cr += wsum(w * r); cg += wsum(w * g); cb += wsum(w * b);
        carry_sd += wlast(incl);
    }
    if (lane == 0) {
        const float one_m = 1.f - W;
        float dist = D, r = cr, g = cg, b = cb;
        a.dacc[ray] = D; a.dl[ray] = du * (1.f / 3.f) + 2.f * db;
        if (a.training) {                                             // nerf_renderer.py:192-194
            float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
            if (a.bg_noise) { n0 = a.bg_noise[4 * ray]; n1 = a.bg_noise[4 * ray + 1]; n2 = a.bg_noise[4 * ray + 2]; n3 = a.bg_noise[4 * ray + 3]; }
            dist = fmaxf(dist + (n3 * 2.f - 1.f) * one_m, 0.f);
            r += n0 * one_m; g += n1 * one_m; b += n2 * one_m;
        } else {                                                      // nerf_renderer.py:195-197
            dist += 5.f * one_m;
            r += 0.5f * one_m; g += 0.5f * one_m; b += 0.5f * one_m;
        }
        a.rgb_out[3 * ray] = r; a.rgb_out[3 * ray + 1] = g; a.rgb_out[3 * ray + 2] = b;
        a.dist_out[ray] = dist; a.op_out[ray] = W;
    }
}

// Backward (same formulas as composite_bwd_ray in train.cu), one warp per ray, chunks from the END of the ray:
// the suffix sums over later samples are carries; inside a chunk suffix_i = chunk_total - inclusive_prefix_i.
template <int PHASE>
__global__ void __launch_bounds__(256) composite_packed_bwd_kernel(const PkCompArgs a)
{
    const int lane = threadIdx.x & 31;
    const uint64_t ray = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= a.R) return;
    const int64_t b0 = a.offsets[ray], b1 = a.offsets[ray + 1];
    if (b1 <= b0) return;
    if constexpr (PHASE == PERF_PHASE_APP) {
        float gr = 0.f, gg = 0.f, gb = 0.f;
        if (a.g_rgb) { gr = a.g_rgb[3 * ray]; gg = a.g_rgb[3 * ray + 1]; gb = a.g_rgb[3 * ray + 2]; }
        for (int64_t n = b0 + lane; n < b1; n += 32) {
            const float w = a.w[n];
            const uint2 c = *reinterpret_cast<const uint2*>(a.rgb + n * 4);
            const float2 c01 = unpack_half2(c.x), c2 = unpack_half2(c.y);
            a.dz[n * 3 + 0] = gr * w * c01.x * (1.f - c01.x);
            a.dz[n * 3 + 1] = gg * w * c01.y * (1.f - c01.y);
            a.dz[n * 3 + 2] = gb * w * c2.x * (1.f - c2.x);
        }
    } else {
        const float O = a.op_out[ray], Dacc = a.dacc[ray];
        float cbg = 0.f;
        if (a.bg_noise) cbg = a.bg_noise[4 * ray + 3] * 2.f - 1.f;
        const float mask = a.dist_out[ray] > 0.f ? 1.f : 0.f;
        const float gd = (a.g_dist ? a.g_dist[ray] : 0.f) * mask;
        const float gO = (a.g_op ? a.g_op[ray] : 0.f) - gd * cbg;
        const float gdl = a.g_dl ? a.g_dl[ray] : 0.f;
        float Wsuf_c = 0.f, WMsuf_c = 0.f, wg_c = 0.f;                   // sums over the chunks already done (later samples)
        const int64_t n_chunks = (b1 - b0 + 31) / 32;
        for (int64_t ch = n_chunks - 1; ch >= 0; --ch) {
            const int64_t n = b0 + ch * 32 + lane;
            const bool live = n < b1;
            float ts = 0.f, te = 0.f, w = 0.f, T = 0.f, sig = 0.f;
            if (live) { ts = a.ts[n]; te = a.te[n]; w = a.w[n]; T = a.T[n]; sig = a.sigma[n]; }
            const float dt = __fsub_rn(te, ts), m = __fadd_rn(ts, te) * 0.5f;
            const float wm = w * m;
            const float wi = wscan(w, lane), wmi = wscan(wm, lane);
            const float wt = wlast(wi), wmt = wlast(wmi);
            const float Wsuf = Wsuf_c + (wt - wi), WMsuf = WMsuf_c + (wmt - wmi);     // over later samples, exclusive
            const float Wx = O - Wsuf - w, WMx = Dacc - WMsuf - wm;
            const float ddl = (2.f / 3.f) * dt * w + 2.f * (m * Wx - WMx) + 2.f * (WMsuf - m * Wsuf);
            const float g = gd * m + gO + gdl * ddl;
            const float wg = w * g;
            const float wgi = wscan(wg, lane);
            const float wgt = wlast(wgi);
            const float suf_wg = wg_c + (wgt - wgi);
            const float dsd = (T - w) * g - suf_wg;
            // trunc_exp backward (ngp_nerf.py:36-38); T == 0 marks a dropped sample: no gradient
            if (live) a.dz[n] = T > 0.f ? dsd * dt * fminf(sig, 3269017.3724721107f) : 0.f;
            Wsuf_c += wt; WMsuf_c += wmt; wg_c += wgt;
        }
    }
}

// ---- occupancy-grid update (nerfacc OccGridEstimator._update, levels = 1; nerf.py:159-168) -----------------------
// splitmix-style counter hash -> U[0,1): jitter of the evaluation point inside its cell
__device__ __forceinline__ float u01(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; x ^= x >> 31;
    return (float)(uint32_t)(x >> 40) * (1.0f / 16777216.0f);
}
// x[i] = aabb_min + (cell coords of idx[i] + U[0,1)^3) / res * extent ; idx == null: cell i
__global__ void __launch_bounds__(256) occ_points_kernel(const int64_t* __restrict__ idx, uint64_t n, int rx, int ry, int rz,
                                                         float ax, float ay, float az, float ex, float ey, float ez,
                                                         uint64_t seed, float* __restrict__ x)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = idx ? idx[i] : (int64_t)i;
    const int cz = (int)(c % rz), cy = (int)((c / rz) % ry), cx = (int)(c / ((int64_t)rz * ry));
    const uint64_t k = seed + 3 * i;
    x[3 * i]     = ax + ((float)cx + u01(k))     / (float)rx * ex;
    x[3 * i + 1] = ay + ((float)cy + u01(k + 1)) / (float)ry * ey;
    x[3 * i + 2] = az + ((float)cz + u01(k + 2)) / (float)rz * ez;
}
// occs[c] = max(occs[c] * decay, occ[i]) for the evaluated cells
__global__ void __launch_bounds__(256) occ_ema_kernel(float* __restrict__ occs, const int64_t* __restrict__ idx, const float* __restrict__ occ,
                                                      uint64_t n, float decay)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = idx ? idx[i] : (int64_t)i;
    occs[c] = fmaxf(occs[c] * decay, occ[i]);
}
// deterministic two-stage mean of occs (only cells >= 0 count, as upstream) + threshold -> binaries
__global__ void __launch_bounds__(256) occ_partial_kernel(const float* __restrict__ occs, uint64_t n, double* __restrict__ part /*[2*grid]*/)
{
    __shared__ double ss[8], sc[8];
    double s = 0.0, c = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float v = occs[i];
        if (v >= 0.f) { s += (double)v; c += 1.0; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); c += __shfl_xor_sync(0xffffffffu, c, off); }
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sc[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { s += ss[w]; c += sc[w]; }
        part[2 * blockIdx.x] = s; part[2 * blockIdx.x + 1] = c;
    }
}
__global__ void __launch_bounds__(256) occ_binarise_kernel(const float* __restrict__ occs, uint64_t n, const double* __restrict__ part, int n_part,
                                                           float occ_thre, uint8_t* __restrict__ binaries)
{
    __shared__ float s_thre;
    if (threadIdx.x == 0) {
        double s = 0.0, c = 0.0;
        for (int i = 0; i < n_part; ++i) { s += part[2 * i]; c += part[2 * i + 1]; }     // fixed order: every block gets the same value
        const float mean = c > 0.0 ? (float)(s / c) : 0.f;
        s_thre = fminf(mean, occ_thre);
    }
    __syncthreads();
    const float thre = s_thre;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        binaries[i] = occs[i] > thre ? 1 : 0;
}

}  // namespace perf

using namespace perf;

static int fill_pk(PkCompArgs& a, const int64_t* off, const float* ts, const float* te, const float* sigma, const void* rgb, uint64_t R)
{
    PERF_CHECK_ARG(off && ts && te && sigma && rgb, "NULL pointer");
    PERF_CHECK_ARG((uintptr_t)rgb % 8 == 0, "misaligned rgb");
    memset(&a, 0, sizeof(a));
    a.offsets = off; a.ts = ts; a.te = te; a.sigma = sigma; a.rgb = (const __half*)rgb; a.R = R;
    return PERF_OK;
}

extern "C" {
#pragma GCC visibility push(default)

int perf_composite_packed_fwd(const int64_t* d_offsets, const float* d_t_starts, const float* d_t_ends, const float* d_sigma,
                              const void* d_rgb_half4, uint64_t R, float early_stop_eps, uint32_t flags, const float* d_bg_noise,
                              float* d_weights, float* d_trans, float* d_rgb_out, float* d_distance_out, float* d_opacity_out,
                              float* d_dist_acc, float* d_distloss, void* stream)
{
    PkCompArgs a; int rc = fill_pk(a, d_offsets, d_t_starts, d_t_ends, d_sigma, d_rgb_half4, R); if (rc) return rc;
    PERF_CHECK_ARG(d_weights && d_trans && d_rgb_out && d_distance_out && d_opacity_out && d_dist_acc && d_distloss, "NULL output");
    a.eps = early_stop_eps; a.training = (flags & PERF_FLAG_TRAINING) ? 1 : 0; a.bg_noise = d_bg_noise;
    a.w = d_weights; a.T = d_trans; a.rgb_out = d_rgb_out; a.dist_out = d_distance_out; a.op_out = d_opacity_out;
    a.dacc = d_dist_acc; a.dl = d_distloss;
    if (R == 0) return PERF_OK;
    composite_packed_fwd_kernel<<<(unsigned)((R * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    PERF_LAUNCH_CHECK();
    return PERF_OK;
}

int perf_composite_packed_bwd(int phase, const int64_t* d_offsets, const float* d_t_starts, const float* d_t_ends, const float* d_sigma,
                              const void* d_rgb_half4, uint64_t R, const float* d_bg_noise, const float* d_weights, const float* d_trans,
                              const float* d_distance_out, const float* d_opacity_out, const float* d_dist_acc,
                              const float* d_g_rgb, const float* d_g_distance, const float* d_g_opacity, const float* d_g_distloss,
                              float* d_dz, void* stream)
{
    PkCompArgs a; int rc = fill_pk(a, d_offsets, d_t_starts, d_t_ends, d_sigma, d_rgb_half4, R); if (rc) return rc;
    PERF_CHECK_ARG(phase == PERF_PHASE_GEO || phase == PERF_PHASE_APP, "bad phase");
    PERF_CHECK_ARG(d_weights && d_trans && d_dz, "NULL pointer");
    PERF_CHECK_ARG(phase == PERF_PHASE_APP || (d_distance_out && d_opacity_out && d_dist_acc), "density phase needs the forward outputs");
    a.bg_noise = d_bg_noise; a.w = const_cast<float*>(d_weights); a.T = const_cast<float*>(d_trans);
    a.dist_out = const_cast<float*>(d_distance_out); a.op_out = const_cast<float*>(d_opacity_out); a.dacc = const_cast<float*>(d_dist_acc);
    a.g_rgb = d_g_rgb; a.g_dist = d_g_distance; a.g_op = d_g_opacity; a.g_dl = d_g_distloss; a.dz = d_dz;
    if (R == 0) return PERF_OK;
    const unsigned grid = (unsigned)((R * 32 + 255) / 256);
    if (phase == PERF_PHASE_GEO) composite_packed_bwd_kernel<PERF_PHASE_GEO><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else composite_packed_bwd_kernel<PERF_PHASE_APP><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    PERF_LAUNCH_CHECK();
    return PERF_OK;
}

int perf_occ_points(const int64_t* d_cell_idx, uint64_t n, const int* h_res3, const float* h_aabb6, uint64_t seed, float* d_x, void* stream)
{
    PERF_CHECK_ARG(h_res3 && h_aabb6 && d_x, "NULL pointer");
    PERF_CHECK_ARG(h_res3[0] > 0 && h_res3[1] > 0 && h_res3[2] > 0, "bad resolution");
    if (n == 0) return PERF_OK;
    occ_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_cell_idx, n, h_res3[0], h_res3[1], h_res3[2],
        h_aabb6[0], h_aabb6[1], h_aabb6[2], h_aabb6[3] - h_aabb6[0], h_aabb6[4] - h_aabb6[1], h_aabb6[5] - h_aabb6[2], seed, d_x);
    PERF_LAUNCH_CHECK();
    return PERF_OK;
}

int perf_occ_update(float* d_occs, uint64_t n_cells, const int64_t* d_cell_idx, const float* d_occ_new, uint64_t n, float ema_decay,
                    float occ_thre, uint8_t* d_binaries, double* d_workspace /* >= 2 * PERF_OCC_PARTIALS doubles */, void* stream)
{
    PERF_CHECK_ARG(d_occs && d_occ_new && d_binaries && d_workspace, "NULL pointer");
    PERF_CHECK_ARG(d_cell_idx != nullptr || n == n_cells, "all-cell update needs n == n_cells");
    if (n_cells == 0) return PERF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (n > 0) { occ_ema_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_occs, d_cell_idx, d_occ_new, n, ema_decay); PERF_LAUNCH_CHECK(); }
    occ_partial_kernel<<<PERF_OCC_PARTIALS, 256, 0, st>>>(d_occs, n_cells, d_workspace);
    PERF_LAUNCH_CHECK();
    occ_binarise_kernel<<<(unsigned)(num_sms() * 8), 256, 0, st>>>(d_occs, n_cells, d_workspace, PERF_OCC_PARTIALS, occ_thre, d_binaries);
    PERF_LAUNCH_CHECK();
    return PERF_OK;
}

#pragma GCC visibility pop
}  // extern "C"
