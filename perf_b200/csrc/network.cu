// network.cu -- tcnn NetworkWithInputEncoding.forward as ONE kernel: hash-grid encode fused with
// the 64-wide MLP on tcgen05 (features never leave the SM).  Also the stand-alone MLP forward.
//
// Per 128-sample tile (thread t = sample t):
//   encode 16 levels (8 half2 gathers each, fp32 blend) -> fp16 features -> smem A tile (K=32)
//   thread 0: 2 x tcgen05.mma (K=16 each) -> D1 in TMEM, tcgen05.commit -> mbarrier
//   all: tcgen05.ld own row, ReLU, round fp16;  1 hidden layer: dot with the output row(s)
//        2 hidden layers: store H tile (K=64), 4 x tcgen05.mma -> D2, ld, ReLU, dots.
#include "mlp_tc.cuh"

namespace perf {

struct NetArgs {
    LevelTable      lt;
    const uint32_t* table;      // fp16x2 entries (ENCODE only)
    const __half*   weights;    // fp16 MLP matrices: W1 [64,32] | (W2 [64,64]) | Wout [16,64]
    const float*    x01;        // [N,3]          (ENCODE)
    const uint4*    feat_in;    // [N,32] fp16    (!ENCODE)
    uint64_t        N;
    uint32_t        n_out, out_act;
    __half*         out;        // [N, n_out]
    uint4*          feat_save;  // [N,32] fp16 or null
    uint4*          h1_save;    // [N,64] fp16 or null
    uint4*          h2_save;    // [N,64] fp16 or null
};

constexpr int NET_SMEM_A    = 0;
constexpr int NET_SMEM_H    = NET_SMEM_A + A32_BYTES;
constexpr int NET_SMEM_W1   = NET_SMEM_H + A64_BYTES;
constexpr int NET_SMEM_W2   = NET_SMEM_W1 + W32_BYTES;
constexpr int NET_SMEM_WOUT = NET_SMEM_W2 + W64_BYTES;
constexpr int NET_SMEM_BAR  = NET_SMEM_WOUT + 16 * HID * 4;
constexpr int NET_SMEM_TOTAL = NET_SMEM_BAR + 16;

template <bool ENCODE, bool TWO_HIDDEN, bool SIMT>
__global__ void __launch_bounds__(TILE, 4) network_fwd_kernel(const __grid_constant__ NetArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem + NET_SMEM_A;
    uint8_t* sH = smem + NET_SMEM_H;
    uint8_t* sW1 = smem + NET_SMEM_W1;
    uint8_t* sW2 = smem + NET_SMEM_W2;
    float*   sWout = reinterpret_cast<float*>(smem + NET_SMEM_WOUT);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NET_SMEM_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + NET_SMEM_BAR + 8);

    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- one-time setup: weights -> smem, mbarrier, TMEM
    load_weight_canonical(a.weights, 32, sW1, tid, TILE);
    const __half* wnext = a.weights + HID * 32;
    if (TWO_HIDDEN) { load_weight_canonical(wnext, 64, sW2, tid, TILE); wnext += HID * HID; }
    load_wout(wnext, a.n_out, sWout, tid, TILE);
    uint32_t tmem_base = 0;
    if (!SIMT) {
        if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
        __syncwarp();
        if (warp == 0) tmem_alloc<64>(tmem_slot);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tmem_base = *tmem_slot;
    } else {
        __syncthreads();
    }
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t parity = 0;

    const uint64_t n_tiles = (a.N + TILE - 1) / TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t i = tile * TILE + tid;
        const bool valid = i < a.N;

        // ---- stage the fp16 feature row
        uint32_t packed[16];
        if constexpr (ENCODE) {
            float x = 0.5f, y = 0.5f, z = 0.5f;
            if (valid) { x = a.x01[3 * i]; y = a.x01[3 * i + 1]; z = a.x01[3 * i + 2]; }
#pragma unroll
            for (int l = 0; l < 16; ++l) {
                Corner8 c; level_corners(a.lt, l, x, y, z, c);
                uint32_t v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = __ldg(a.table + c.idx[k]);
                packed[l] = blend8_half(c.w, v);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 u = valid ? a.feat_in[i * 4 + q] : make_uint4(0, 0, 0, 0);
                packed[4 * q] = u.x; packed[4 * q + 1] = u.y; packed[4 * q + 2] = u.z; packed[4 * q + 3] = u.w;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 u = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            *reinterpret_cast<uint4*>(sA + (q * TILE + tid) * 16) = u;
            if (a.feat_save && valid) a.feat_save[i * 4 + q] = u;
        }

        // ---- layer 1
        if constexpr (!SIMT) {
            fence_proxy_async();       // my st.shared -> visible to the async (tensor-core) proxy
            tc_fence_before();         // my previous tcgen05.ld of D is ordered before the barrier
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                issue_layer(tmem_base, smem_u32(sA), smem_u32(sW1), 32);
                umma_commit(bar);
            }
            mbar_wait(bar, parity); parity ^= 1u;
            tc_fence_after();
        } else {
            __syncthreads();
        }

        float2 out_acc[16];                     // even / odd partial sums of every output (mlp_tc.cuh::out_dots)
#pragma unroll
        for (int o = 0; o < 16; ++o) out_acc[o] = make_float2(0.f, 0.f);

#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            float v[32]; uint32_t hp[16];
            acc_chunk<SIMT>(c, 32, tmem_row, sA, sW1, tid, v);
            relu_pack(v, hp);
            if (a.h1_save && valid) store_chunk_global(a.h1_save + i * 8, c, hp);
            if constexpr (TWO_HIDDEN) store_chunk_canonical(sH, tid, 4 * c, hp);
            else out_dots<16>(hp, sWout, c, (int)a.n_out, out_acc);
        }

        if constexpr (TWO_HIDDEN) {
            // ---- layer 2
            if constexpr (!SIMT) {
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    issue_layer(tmem_base, smem_u32(sH), smem_u32(sW2), 64);
                    umma_commit(bar);
                }
                mbar_wait(bar, parity); parity ^= 1u;
                tc_fence_after();
            } else {
                __syncthreads();
            }
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float v[32]; uint32_t hp[16];
                acc_chunk<SIMT>(c, 64, tmem_row, sH, sW2, tid, v);
                relu_pack(v, hp);
                if (a.h2_save && valid) store_chunk_global(a.h2_save + i * 8, c, hp);
                out_dots<16>(hp, sWout, c, (int)a.n_out, out_acc);
            }
        }

        if (valid) {
#pragma unroll
            for (int o = 0; o < 16; ++o)
                if (o < (int)a.n_out) a.out[i * a.n_out + o] = __float2half_rn(finish_output(out_sum(out_acc[o]), a.out_act));
        }
        if constexpr (SIMT) __syncthreads();   // rows of sA/sH are rewritten next tile
    }

    if (!SIMT) {
        tc_fence_before();
        __syncthreads();
        if (warp == 0) tmem_dealloc<64>(tmem_base);
    }
}

template <bool ENCODE>
static int launch_network(const NetArgs& a, bool two_hidden, bool simt, cudaStream_t stream)
{
    const uint64_t n_tiles = (a.N + TILE - 1) / TILE;
    const unsigned grid = (unsigned)(n_tiles < (uint64_t)num_sms() * 4 ? n_tiles : (uint64_t)num_sms() * 4);
#define PERF_NET_LAUNCH(TH, SM) do { \
        auto k = network_fwd_kernel<ENCODE, TH, SM>; \
        PERF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, NET_SMEM_TOTAL)); \
        k<<<grid, TILE, NET_SMEM_TOTAL, stream>>>(a); } while (0)
    if (two_hidden) { if (simt) PERF_NET_LAUNCH(true, true); else PERF_NET_LAUNCH(true, false); }
    else            { if (simt) PERF_NET_LAUNCH(false, true); else PERF_NET_LAUNCH(false, false); }
#undef PERF_NET_LAUNCH
    PERF_LAUNCH_CHECK();
    return PERF_OK;
}

}  // namespace perf

using namespace perf;

extern "C" {
#pragma GCC visibility push(default)

int perf_network_fwd(const perf_grid_cfg* grid, const perf_mlp_cfg* mlp, const void* d_params_half,
                     const float* d_x01, uint64_t N, void* d_out, void* d_feat, void* d_h1, void* d_h2,
                     uint32_t flags, void* stream)
{
    PERF_CHECK_ARG(d_params_half && d_x01 && d_out, "NULL pointer");
    NetArgs a; memset(&a, 0, sizeof(a));
    int rc = build_level_table(grid, &a.lt, nullptr); if (rc) return rc;
    PERF_CHECK_SUP(grid->n_levels == 16, "fused network kernel needs n_levels == 16 (got %u)", grid->n_levels);
    uint64_t nm = 0; rc = mlp_param_count(mlp, &nm); if (rc) return rc;
    PERF_CHECK_ARG((uintptr_t)d_params_half % 16 == 0, "params_half must be 16-byte aligned");
    PERF_CHECK_ARG((!d_feat || (uintptr_t)d_feat % 16 == 0) && (!d_h1 || (uintptr_t)d_h1 % 16 == 0) && (!d_h2 || (uintptr_t)d_h2 % 16 == 0), "save buffers must be 16-byte aligned");
    if (N == 0) return PERF_OK;
    a.weights = (const __half*)d_params_half;
    a.table = reinterpret_cast<const uint32_t*>((const __half*)d_params_half + nm);
    a.x01 = d_x01; a.N = N; a.n_out = mlp->n_out; a.out_act = mlp->output_activation;
    a.out = (__half*)d_out; a.feat_save = (uint4*)d_feat; a.h1_save = (uint4*)d_h1;
    a.h2_save = mlp->n_hidden_layers == 2 ? (uint4*)d_h2 : nullptr;
    return launch_network<true>(a, mlp->n_hidden_layers == 2, (flags & PERF_FLAG_SIMT_MLP) != 0, (cudaStream_t)stream);
}

int perf_mlp_fwd(const perf_mlp_cfg* mlp, const void* d_weights_half, const void* d_in, uint64_t N,
                 void* d_out, void* d_h1, void* d_h2, uint32_t flags, void* stream)
{
    PERF_CHECK_ARG(d_weights_half && d_in && d_out, "NULL pointer");
    int rc = check_mlp(mlp); if (rc) return rc;
    PERF_CHECK_ARG((uintptr_t)d_weights_half % 16 == 0 && (uintptr_t)d_in % 16 == 0, "weights / input must be 16-byte aligned");
    if (N == 0) return PERF_OK;
    NetArgs a; memset(&a, 0, sizeof(a));
    a.weights = (const __half*)d_weights_half; a.feat_in = (const uint4*)d_in; a.N = N;
    a.n_out = mlp->n_out; a.out_act = mlp->output_activation; a.out = (__half*)d_out;
    a.h1_save = (uint4*)d_h1; a.h2_save = mlp->n_hidden_layers == 2 ? (uint4*)d_h2 : nullptr;
    return launch_network<false>(a, mlp->n_hidden_layers == 2, (flags & PERF_FLAG_SIMT_MLP) != 0, (cudaStream_t)stream);
}

#pragma GCC visibility pop
}  // extern "C"
