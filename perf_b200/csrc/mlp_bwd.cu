// mlp_bwd.cu -- backward of the 64-wide bias-free MLP from the saved fp16 activations, ONE kernel on the
// 5th-gen tensor cores (replaces tcnn FullyFusedMLP::backward_impl for PeRF's two networks; today's default
// path hands the five matrix products to cuBLAS from Python, ops.mlp_backward_half).
//
// STATUS: written at the end of round 1 WITHOUT GPU time left to run it; it is NOT on any default path
// (ops.mlp_backward_half only calls it when PERF_B200_TC_MLP_BWD=1, its GPU tests are gated by
// PERF_B200_EXPERIMENTAL=1).  It carries a CUDA-core twin (PERF_FLAG_SIMT_MLP) that reads the same
// shared-memory operand images, so descriptor mistakes can be told from arithmetic ones.
//
// Rows are samples (any order), thread t of a 128-thread CTA owns row t of the current 128-row tile.
//   colour net (two hidden layers), per tile:
//     dh2 = (dz Wout) . [h2 > 0]                    CUDA cores (n_out = 3)
//     D1  = dh2 W2          -> dh1 = D1 . [h1 > 0]  tcgen05, A K-major, B = forward W2 image read MN-major
//     D2  = dh1 W1          -> dfeat                tcgen05, same
//     G1 += [dh2|dh1]^T [h1|feat]                   tcgen05, A and B MN-major: rows 0-63 x cols 0-63 = dW2,
//                                                   rows 64-127 x cols 64-95 = dW1 (the other blocks are unused)
//     G2 += [h2|h1]^T [dz|0]                        rows 0-63 x cols 0-2 = dWout^T
//   density net (one hidden layer): dh1 = (dz Wout) . [h1 > 0];  D2 = dh1 W1;  G += [dh1|h1]^T [feat|dz|0]
//     (rows 0-63 x cols 0-31 = dW1, rows 64-127 x col 32 = dWout).
// The weight-gradient accumulators G live in TMEM for the whole kernel (accumulate flag) and are added to the
// global fp32 gradient once per CTA.  All operand images use the forward's no-swizzle K-major canonical layout
// (mlp_tc.cuh); "transposed" operands are the SAME images described as MN-major:
//     MN-major, no swizzle:  8 contiguous MN elements (16 B), 8 K-rows at 16 B stride, LBO = byte distance
//     between 8-row K groups (128), SBO = byte distance between 8-element MN groups (rows * 16)
// (cute::UMMA canonical layout ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO))), with bits 15 / 16 of the instruction
// descriptor selecting MN-major for A / B.
#include "mlp_tc.cuh"

namespace perf {

struct MlpBwdArgs {
    const __half* w;         // flat fp16 MLP params (W1 [64,32] | W2 [64,64] (two hidden) | Wout [16,64])
    const uint4*  feat;      // [N,32] fp16
    const uint4*  h1;        // [N,64] fp16
    const uint4*  h2;        // [N,64] fp16 (two hidden) or NULL
    const float*  dz;        // [N,n_out] fp32: d loss / d output pre-activation
    uint64_t      N;
    float*        dW;        // flat fp32 gradient of the MLP params, accumulated (+=)
    float*        dfeat;     // [N,32] fp32
    uint32_t      n_out;     // 1..3
};

__host__ __device__ constexpr uint32_t idesc_f16_major(int M, int N, bool a_mn, bool b_mn)
{
    return umma_idesc_f16(M, N) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16);
}
// MN-major view of a canonical image with `rows` rows, K-step ks (16 rows)
__device__ __forceinline__ uint64_t desc_mn(uint32_t img, int rows, int ks) { return umma_desc(img + ks * 256, 128u, (uint32_t)rows * 16u); }
// K-major view of an activation image (128 rows), K-step ks (16 columns)
__device__ __forceinline__ uint64_t desc_k(uint32_t img, int ks) { return umma_desc(img + ks * 2 * A_LBO, A_LBO, X_SBO); }

__device__ __forceinline__ float img_at(const uint8_t* img, int rows, int r, int c)
{
    return __half2float(*reinterpret_cast<const __half*>(img + ((c >> 3) * rows + r) * 16 + (c & 7) * 2));
}

// shared-memory map (bytes)
template <bool TWO> struct BwdSmem;
template <> struct BwdSmem<true> {
    static constexpr int DH2 = 0, DH1 = DH2 + A64_BYTES, H2 = DH1 + A64_BYTES, H1 = H2 + A64_BYTES, FEAT = H1 + A64_BYTES,
                         DZP = FEAT + A32_BYTES, W2 = DZP + 2 * TILE * 16, W1 = W2 + W64_BYTES, WOUT = W1 + W32_BYTES,
                         BAR = WOUT + 3 * HID * 4, TOTAL = BAR + 16;
    static constexpr int TM_D1 = 0, TM_D2 = 64, TM_G1 = 96, TM_G2 = 192, TM_COLS = 256, N_G1 = 96, N_G2 = 16;
};
template <> struct BwdSmem<false> {
    static constexpr int DH1 = 0, H1 = DH1 + A64_BYTES, FEAT = H1 + A64_BYTES, DZP = FEAT + A32_BYTES, W1 = DZP + 2 * TILE * 16,
                         WOUT = W1 + W32_BYTES, BAR = WOUT + 3 * HID * 4, TOTAL = BAR + 16;
    static constexpr int TM_D2 = 0, TM_G1 = 32, TM_COLS = 128, N_G1 = 48;
};

template <int KGS>
__device__ __forceinline__ void store_row(uint8_t* img, int row, const uint4 (&v)[KGS])
{
#pragma unroll
    for (int kg = 0; kg < KGS; ++kg) *reinterpret_cast<uint4*>(img + (kg * TILE + row) * 16) = v[kg];
}

// dh[j] = (h[j] > 0) ? sum_o dz[o] * wout[o][j] : 0, rounded to fp16, written as this thread's row of `dst`
__device__ __forceinline__ void out_layer_backward(uint8_t* dst, int row, const uint4 (&h)[8], const float (&dz)[3], int n_out, const float* wout)
{
#pragma unroll
    for (int kg = 0; kg < 8; ++kg) {
        const uint32_t hw[4] = {h[kg].x, h[kg].y, h[kg].z, h[kg].w};
        uint32_t o4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 hh = unpack_half2(hw[q]);
            const int j = kg * 8 + 2 * q;
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int o = 0; o < 3; ++o)
                if (o < n_out) { a = fmaf(dz[o], wout[o * HID + j], a); b = fmaf(dz[o], wout[o * HID + j + 1], b); }
            o4[q] = pack_half2(hh.x > 0.f ? a : 0.f, hh.y > 0.f ? b : 0.f);
        }
        *reinterpret_cast<uint4*>(dst + (kg * TILE + row) * 16) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
    }
}

// CUDA-core twins of the MMAs, reading the same images.
// D[row][n] = sum_k A[row][k] * Wimg[k][n]   (A: activation image K-major; Wimg: forward weight image [64 rows k][ncols])
__device__ __forceinline__ void simt_dgrad(const uint8_t* A, const uint8_t* Wimg, int row, int n0, float (&v)[32])
{
    for (int j = 0; j < 32; ++j) {
        float acc = 0.f;
        for (int k = 0; k < HID; ++k) acc = fmaf(img_at(A, TILE, row, k), img_at(Wimg, HID, k, n0 + j), acc);
        v[j] = acc;
    }
}

template <bool TWO, bool SIMT>
__global__ void __launch_bounds__(128) mlp_bwd_kernel(const MlpBwdArgs a)
{
    using L = BwdSmem<TWO>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::BAR + 8);
    float* s_wout = reinterpret_cast<float*>(smem + L::WOUT);
    const int t = threadIdx.x, warp = t >> 5;
    const int n_out = (int)a.n_out;

    // weights: forward canonical images (read MN-major by the dgrad MMAs) + fp32 output matrix
    load_weight_canonical(a.w, 32, smem + L::W1, t, TILE);
    const __half* wout_g = a.w + 64 * 32;
    if constexpr (TWO) { load_weight_canonical(a.w + 64 * 32, 64, smem + L::W2, t, TILE); wout_g += 64 * 64; }
    load_wout(wout_g, n_out, s_wout, t, TILE);
    uint32_t tmem_base = 0;
    if constexpr (!SIMT) {
        if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
        if (warp == 0) tmem_alloc<L::TM_COLS>(tmem_slot);
        tc_fence_before();
    }
    __syncthreads();
    if constexpr (!SIMT) { tc_fence_after(); tmem_base = *tmem_slot; }
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;
    float gacc[SIMT ? 112 : 1];                       // SIMT twin of the TMEM weight-gradient accumulators (row m = t)
    if constexpr (SIMT) { for (int i = 0; i < 112; ++i) gacc[i] = 0.f; }

    const uint64_t n_tiles = (a.N + TILE - 1) / TILE;
    bool first = true;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first = false) {
        const uint64_t row = tile * TILE + t;
        const bool valid = row < a.N;
        uint4 f4[4], h1v[8], hlast[8];
        float dz[3] = {0.f, 0.f, 0.f};
        const uint4 z4 = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int q = 0; q < 4; ++q) f4[q] = valid ? a.feat[row * 4 + q] : z4;
#pragma unroll
        for (int q = 0; q < 8; ++q) h1v[q] = valid ? a.h1[row * 8 + q] : z4;
        if constexpr (TWO) {
#pragma unroll
            for (int q = 0; q < 8; ++q) hlast[q] = valid ? a.h2[row * 8 + q] : z4;
        }
#pragma unroll
        for (int o = 0; o < 3; ++o) if (valid && o < n_out) dz[o] = a.dz[row * n_out + o];

        store_row(smem + L::FEAT, t, f4);
        store_row(smem + L::H1, t, h1v);
        {   // dz as 16 fp16 columns (cols >= n_out are zero)
            uint8_t* p = smem + L::DZP;
            *reinterpret_cast<uint4*>(p + (0 * TILE + t) * 16) = make_uint4(pack_half2(dz[0], dz[1]), pack_half2(dz[2], 0.f), 0u, 0u);
            *reinterpret_cast<uint4*>(p + (1 * TILE + t) * 16) = z4;
        }
        if constexpr (TWO) {
            store_row(smem + L::H2, t, hlast);
            out_layer_backward(smem + L::DH2, t, hlast, dz, n_out, s_wout);
        } else {
            out_layer_backward(smem + L::DH1, t, h1v, dz, n_out, s_wout);
        }
        if constexpr (!SIMT) { fence_proxy_async(); tc_fence_before(); }
        __syncthreads();

        if constexpr (TWO) {
            // ---- D1 = dh2 W2 -> dh1 = D1 . [h1 > 0]
            if constexpr (!SIMT) {
                if (t == 0) {
                    tc_fence_after();
                    constexpr uint32_t id = idesc_f16_major(TILE, HID, false, true);
                    for (int ks = 0; ks < 4; ++ks)
                        umma_f16(tmem_base + L::TM_D1, desc_k(smem_u32(smem + L::DH2), ks), desc_mn(smem_u32(smem + L::W2), HID, ks), id, ks > 0);
                    umma_commit(bar);
                }
                mbar_wait(bar, phase); phase ^= 1u;
                tc_fence_after();
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float v[32];
                if constexpr (SIMT) simt_dgrad(smem + L::DH2, smem + L::W2, t, 32 * c, v);
                else tmem_ld32(tmem_row + L::TM_D1 + 32 * c, v);
                uint32_t p[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint4 hq = h1v[4 * c + j / 4];
                    const float2 hh = unpack_half2((j % 4) == 0 ? hq.x : (j % 4) == 1 ? hq.y : (j % 4) == 2 ? hq.z : hq.w);
                    p[j] = pack_half2(hh.x > 0.f ? v[2 * j] : 0.f, hh.y > 0.f ? v[2 * j + 1] : 0.f);
                }
                store_chunk_canonical(smem + L::DH1, t, 4 * c, p);
            }
            if constexpr (!SIMT) { fence_proxy_async(); tc_fence_before(); }
            __syncthreads();
        }

        // ---- D2 = dh1 W1 (-> dfeat) and the weight-gradient accumulations
        if constexpr (!SIMT) {
            if (t == 0) {
                tc_fence_after();
                constexpr uint32_t id2 = idesc_f16_major(TILE, 32, false, true);
                for (int ks = 0; ks < 4; ++ks)
                    umma_f16(tmem_base + L::TM_D2, desc_k(smem_u32(smem + L::DH1), ks), desc_mn(smem_u32(smem + L::W1), HID, ks), id2, ks > 0);
                constexpr uint32_t idg = idesc_f16_major(TILE, L::N_G1, true, true);
                const uint32_t a_img = smem_u32(smem + (TWO ? 0 /* DH2|DH1 */ : L::DH1 /* DH1|H1 */));
                const uint32_t b_img = smem_u32(smem + (TWO ? L::H1 /* H1|FEAT */ : L::FEAT /* FEAT|DZP */));
                for (int ks = 0; ks < 8; ++ks)
                    umma_f16(tmem_base + L::TM_G1, desc_mn(a_img, TILE, ks), desc_mn(b_img, TILE, ks), idg, (!first || ks > 0) ? 1u : 0u);
                if constexpr (TWO) {
                    constexpr uint32_t idg2 = idesc_f16_major(TILE, BwdSmem<true>::N_G2, true, true);
                    for (int ks = 0; ks < 8; ++ks)
                        umma_f16(tmem_base + BwdSmem<true>::TM_G2, desc_mn(smem_u32(smem + BwdSmem<true>::H2), TILE, ks),
                                 desc_mn(smem_u32(smem + L::DZP), TILE, ks), idg2, (!first || ks > 0) ? 1u : 0u);
                }
                umma_commit(bar);
            }
            mbar_wait(bar, phase); phase ^= 1u;
            tc_fence_after();
        } else {
            // weight gradients, row m = t of G: sum over the tile's 128 rows of A[n][m] * B[n][c]
            const uint8_t* a_img = smem + (TWO ? 0 : L::DH1);
            const uint8_t* b_img = smem + (TWO ? L::H1 : L::FEAT);
            for (int c = 0; c < L::N_G1; ++c) {
                float acc = gacc[c];
                for (int n = 0; n < TILE; ++n) acc = fmaf(img_at(a_img, TILE, n, t), img_at(b_img, TILE, n, c), acc);
                gacc[c] = acc;
            }
            if constexpr (TWO) {
                for (int c = 0; c < 16; ++c) {
                    float acc = gacc[96 + c];
                    for (int n = 0; n < TILE; ++n) acc = fmaf(img_at(smem + BwdSmem<true>::H2, TILE, n, t), img_at(smem + L::DZP, TILE, n, c), acc);
                    gacc[96 + c] = acc;
                }
            }
        }
        {
            float v[32];
            if constexpr (SIMT) simt_dgrad(smem + L::DH1, smem + L::W1, t, 0, v);
            else tmem_ld32(tmem_row + L::TM_D2, v);
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(a.dfeat + row * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
        }
        if constexpr (!SIMT) tc_fence_before();
        __syncthreads();                                   // images and D1 / D2 are rewritten by the next tile
    }

    // ---- weight gradients of this CTA -> global (row m = t of the accumulators)
    if (!first) {
        float* dW1 = a.dW;
        float* dW2 = a.dW + 64 * 32;
        float* dWo = a.dW + 64 * 32 + (TWO ? 64 * 64 : 0);
        if constexpr (!SIMT) tc_fence_after();
        for (int c0 = 0; c0 < L::N_G1; c0 += 32) {
            float v[32];
            if constexpr (SIMT) { for (int j = 0; j < 32; ++j) v[j] = (c0 + j < L::N_G1) ? gacc[c0 + j] : 0.f; }
            else tmem_ld32(tmem_row + L::TM_G1 + c0, v);
            for (int j = 0; j < 32; ++j) {
                const int c = c0 + j;
                if (c >= L::N_G1) break;
                if constexpr (TWO) {
                    if (t < 64 && c < 64) atomicAdd(dW2 + t * 64 + c, v[j]);
                    else if (t >= 64 && c >= 64) atomicAdd(dW1 + (t - 64) * 32 + (c - 64), v[j]);
                } else {
                    if (t < 64 && c < 32) atomicAdd(dW1 + t * 32 + c, v[j]);
                    else if (t >= 64 && c >= 32 && c - 32 < n_out) atomicAdd(dWo + (c - 32) * 64 + (t - 64), v[j]);
                }
            }
        }
        if constexpr (TWO) {
            float v[32];
            if constexpr (SIMT) { for (int j = 0; j < 32; ++j) v[j] = j < 16 ? gacc[96 + j] : 0.f; }
            else tmem_ld32(tmem_row + BwdSmem<true>::TM_G2, v);
#pragma unroll
            for (int o = 0; o < 3; ++o) if (t < 64 && o < n_out) atomicAdd(dWo + o * 64 + t, v[o]);
        }
    }
    if constexpr (!SIMT) {
        tc_fence_before();
        __syncthreads();
        if (warp == 0) tmem_dealloc<L::TM_COLS>(tmem_base);
    }
}

template <bool TWO, bool SIMT>
static int launch_mlp_bwd(const MlpBwdArgs& a, cudaStream_t stream)
{
    auto k = mlp_bwd_kernel<TWO, SIMT>;
    static thread_local int attr_dev = -1;
    int dev = 0; PERF_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) { PERF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<TWO>::TOTAL)); attr_dev = dev; }
    const uint64_t n_tiles = (a.N + TILE - 1) / TILE;
    const uint64_t slots = (uint64_t)num_sms() * 2;
    k<<<(unsigned)(n_tiles < slots ? n_tiles : slots), TILE, BwdSmem<TWO>::TOTAL, stream>>>(a);
    PERF_LAUNCH_CHECK();
    return PERF_OK;
}

}  // namespace perf

using namespace perf;

extern "C" {
#pragma GCC visibility push(default)

int perf_mlp_bwd(const perf_mlp_cfg* mlp, const void* d_weights_half, const void* d_feat, const void* d_h1, const void* d_h2,
                 const float* d_dz, uint64_t N, float* d_dweights, float* d_dfeat, uint32_t flags, void* stream)
{
    int rc = check_mlp(mlp); if (rc) return rc;
    PERF_CHECK_ARG(d_weights_half && d_feat && d_h1 && d_dz && d_dweights && d_dfeat, "NULL pointer");
    PERF_CHECK_ARG(mlp->n_hidden_layers == 1 || d_h2, "two hidden layers need d_h2");
    PERF_CHECK_SUP(mlp->n_out <= 3, "n_out=%u (the backward kernel implements 1..3 outputs)", mlp->n_out);
    PERF_CHECK_ARG(((uintptr_t)d_weights_half | (uintptr_t)d_feat | (uintptr_t)d_h1 | (uintptr_t)d_h2 | (uintptr_t)d_dfeat) % 16 == 0, "misaligned buffer");
    if (N == 0) return PERF_OK;
    MlpBwdArgs a;
    a.w = (const __half*)d_weights_half; a.feat = (const uint4*)d_feat; a.h1 = (const uint4*)d_h1; a.h2 = (const uint4*)d_h2;
    a.dz = d_dz; a.N = N; a.dW = d_dweights; a.dfeat = d_dfeat; a.n_out = mlp->n_out;
    const bool simt = (flags & PERF_FLAG_SIMT_MLP) != 0;
    if (mlp->n_hidden_layers == 2) return simt ? launch_mlp_bwd<true, true>(a, (cudaStream_t)stream) : launch_mlp_bwd<true, false>(a, (cudaStream_t)stream);
    return simt ? launch_mlp_bwd<false, true>(a, (cudaStream_t)stream) : launch_mlp_bwd<false, false>(a, (cudaStream_t)stream);
}

#pragma GCC visibility pop
}  // extern "C"
