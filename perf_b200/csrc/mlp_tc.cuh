// mlp_tc.cuh -- building blocks of the 64-wide bias-free MLP on the 5th-gen tensor cores.
//
// One CTA = 128 threads = one 128-row tile (UMMA M=128, cta_group::1); thread t owns row t:
// it produces the row's fp16 input features, later reads the row's fp32 accumulators back from
// TMEM lane t (tcgen05.ld 32x32b: warp w reads lanes 32w..32w+31) and runs the epilogue.
//
// Operand layout in shared memory (both A = activations and B = weights): K-major, NO swizzle,
// i.e. the canonical "interleaved" layout of 8x16-byte core matrices:
//     element (row r, col k)  ->  byte  (k/8) * LBO + r * 16 + (k%8) * 2
// with LBO = rows*16 (all rows of one 8-column k-group are contiguous) and SBO = 128 (8 rows).
// A thread therefore writes its row as K/8 16-byte chunks at stride LBO: consecutive threads
// hit consecutive 16-byte slots -> conflict-free STS.128, no swizzle arithmetic.
#pragma once
#include "common.cuh"

namespace perf {

constexpr int      TILE  = 128;          // rows per tile (UMMA M)
constexpr int      HID   = 64;           // hidden width (UMMA N)
constexpr uint32_t A_LBO = TILE * 16;    // 2048 B between k-groups of an activation tile
constexpr uint32_t W_LBO = HID * 16;     // 1024 B between k-groups of a weight matrix
constexpr uint32_t X_SBO = 128;          // 8 rows * 16 B

constexpr int A32_BYTES = 4 * TILE * 16; //  8 KB : 128 x 32 fp16
constexpr int A64_BYTES = 8 * TILE * 16; // 16 KB : 128 x 64 fp16
constexpr int W32_BYTES = 4 * HID * 16;  //  4 KB :  64 x 32 fp16
constexpr int W64_BYTES = 8 * HID * 16;  //  8 KB :  64 x 64 fp16

// [64, K] row-major fp16 weights in global memory -> canonical K-major smem layout.
__device__ __forceinline__ void load_weight_canonical(const __half* __restrict__ gW, int K, uint8_t* dst, int tid, int nthreads)
{
    const int kgs = K / 8;
    for (int c = tid; c < HID * kgs; c += nthreads) {
        const int n = c % HID, kg = c / HID;
        const uint4 v = *reinterpret_cast<const uint4*>(gW + (size_t)n * K + kg * 8);
        *reinterpret_cast<uint4*>(dst + (kg * HID + n) * 16) = v;
    }
}
// last matrix [16 (padded), 64] fp16 -> fp32 [n_out][64] (fp16 -> fp32 is exact)
__device__ __forceinline__ void load_wout(const __half* __restrict__ gW, int n_out, float* dst, int tid, int nthreads)
{
    for (int c = tid; c < n_out * HID; c += nthreads) dst[c] = __half2float(gW[c]);
}

// One layer: D[128 x 64] (TMEM, fp32) = A[128 x K] * W[64 x K]^T; K/16 MMAs + commit.  ONE thread.
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_smem, uint32_t w_smem, int K)
{
    constexpr uint32_t idesc = umma_idesc_f16(TILE, HID);
    for (int ks = 0; ks < K / 16; ++ks)
        umma_f16(tmem_d, umma_desc(a_smem + ks * 2 * A_LBO, A_LBO, X_SBO),
                         umma_desc(w_smem + ks * 2 * W_LBO, W_LBO, X_SBO), idesc, ks > 0 ? 1u : 0u);
}

__device__ __forceinline__ float dot8(uint4 a, uint4 w, float acc)
{
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 fa = unpack_half2(av[q]), fw = unpack_half2(wv[q]);
        acc = fmaf(fa.x, fw.x, acc); acc = fmaf(fa.y, fw.y, acc);
    }
    return acc;
}

// 32 accumulator columns [32c, 32c+32) of this thread's row.
// TC: from TMEM.  SIMT (debug path, PERF_FLAG_SIMT_MLP): recomputed on CUDA cores from the very
// same shared-memory operands the tensor core would read, so a layout bug shows up as a
// TC-vs-SIMT mismatch.
template <bool SIMT>
__device__ __forceinline__ void acc_chunk(int c, int K, uint32_t tmem_row, const uint8_t* A, const uint8_t* W, int row, float (&v)[32])
{
    if constexpr (SIMT) {
#pragma unroll 1
        for (int j = 0; j < 32; ++j) {
            const int n = 32 * c + j;
            float acc = 0.f;
            for (int kg = 0; kg < K / 8; ++kg)
                acc = dot8(*reinterpret_cast<const uint4*>(A + (kg * TILE + row) * 16),
                           *reinterpret_cast<const uint4*>(W + (kg * HID + n) * 16), acc);
            v[j] = acc;
        }
    } else {
        tmem_ld32(tmem_row + 32 * c, v);
    }
}

// ReLU + round to fp16 (tcnn keeps hidden activations in __half), two values per instruction:
// cvt.rn.relu.f16x2.f32 rounds a pair and clamps it at zero in ONE instruction (round(max(x,0)) == max(round(x),0);
// until round 2 this was F2FP + HMNMX2, 96 more instructions per sample in the render kernel).
__device__ __forceinline__ void relu_pack(const float (&v)[32], uint32_t (&p)[16])
{
#pragma unroll
    for (int j = 0; j < 16; ++j)
        asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(p[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));     // d = {hi: a, lo: b}
}

// Packed fp32 FMA of Blackwell (FFMA2): acc.x += a.x * b.x, acc.y += a.y * b.y in one issue slot.
__device__ __forceinline__ void ffma2(float2& acc, float2 a, float2 b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(reinterpret_cast<unsigned long long&>(acc))
        : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}

// write 32 packed fp16 values as k-groups [kg0, kg0+4) of this thread's row of an activation tile
__device__ __forceinline__ void store_chunk_canonical(uint8_t* A, int row, int kg0, const uint32_t (&p)[16])
{
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(A + ((kg0 + q) * TILE + row) * 16) = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
}
// same 32 values to a row-major [N,64] fp16 global buffer (activation save for the backward pass)
__device__ __forceinline__ void store_chunk_global(uint4* row_ptr /* 8 uint4 per row */, int c, const uint32_t (&p)[16])
{
#pragma unroll
    for (int q = 0; q < 4; ++q)
        row_ptr[4 * c + q] = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
}

// Output layer on CUDA cores (n_out is 1 or 3: a padded N=16 MMA + another smem round trip would cost more than the
// 64*n_out FMAs per row).  Every output keeps TWO partial sums -- acc[o].x over the even hidden units, acc[o].y over
// the odd ones, in increasing order -- so that one FFMA2 serves a packed pair of activations; the pre-activation is
// acc[o].x + acc[o].y (out_sum).  All forward kernels share this order.
//   acc[o] += h[32c + 2j, 32c + 2j + 1] * wout[o][32c + 2j, 32c + 2j + 1]      (wout: fp32 in shared memory)
template <int NOUT_MAX>
__device__ __forceinline__ void out_dots(const uint32_t (&p)[16], const float* wout, int c, int n_out, float2 (&acc)[NOUT_MAX])
{
    float2 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = unpack_half2(p[j]);
#pragma unroll
    for (int o = 0; o < NOUT_MAX; ++o) {
        if (o < n_out) {
            const float4* w4 = reinterpret_cast<const float4*>(wout + o * HID + 32 * c);
            float2 a = acc[o];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 w = w4[q];
                ffma2(a, v[2 * q], make_float2(w.x, w.y));
                ffma2(a, v[2 * q + 1], make_float2(w.z, w.w));
            }
            acc[o] = a;
        }
    }
}
__device__ __forceinline__ float out_sum(float2 acc) { return __fadd_rn(acc.x, acc.y); }

// tcnn output: pre-activation rounded to fp16, activation in fp32, result rounded to fp16
__device__ __forceinline__ float finish_output(float acc, uint32_t out_act)
{
    float o = round_half(acc);
    if (out_act == 1) o = round_half(1.0f / (1.0f + expf(-o)));
    return o;
}

}  // namespace perf
