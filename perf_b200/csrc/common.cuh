// common.cuh -- shared host/device helpers of libperfb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include "../../include/perfb200.h"

namespace perf {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
#define PERF_CHECK_ARG(cond, ...)  do { if (!(cond)) { perf::set_error(__VA_ARGS__); return PERF_EINVAL; } } while (0)
#define PERF_CHECK_SUP(cond, ...)  do { if (!(cond)) { perf::set_error(__VA_ARGS__); return PERF_EUNSUPPORTED; } } while (0)
#define PERF_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
    perf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); return PERF_ECUDA; } } while (0)
#define PERF_LAUNCH_CHECK() PERF_CUDA(cudaGetLastError())

int num_sms();   // cached multiprocessor count of the current device

// ---------------------------------------------------------------- grid level table
// Passed BY VALUE as a kernel parameter (272 bytes) -> lives in the constant bank.
struct LevelTable {
    float    scale[PERF_MAX_LEVELS];
    uint32_t res[PERF_MAX_LEVELS];
    uint32_t size[PERF_MAX_LEVELS];
    uint32_t offset[PERF_MAX_LEVELS];
    uint32_t n_levels;
    uint32_t hashed_mask;      // bit l set: level l is hashed
    uint32_t pow2_mask;        // bit l set: size[l] is a power of two (mod == and)
    uint32_t smoothstep;
};
int build_level_table(const perf_grid_cfg* cfg, LevelTable* out, uint64_t* n_entries);

// Layout of the packed gather table (perf_pack_tables): entries [0, n_entries) in parameter order, then a CELL-MAJOR copy
// of the leading dense levels -- for level l, res_l^3 cells x 8 corners x 8 bytes, cell (gx,gy,gz) at gx + res (gy + res gz),
// corner k (bit0 = x, bit1 = y, bit2 = z) holding entry ((gx+kx) + res (gy+ky) + res^2 (gz+kz)) % size of that level.
// The fused field kernels read a dense level as ONE 64-byte record per sample (4 x LDG.128, one address, no wrap test)
// instead of eight 8-byte gathers.  cell_start[l]: first entry (8-byte units) of level l's cells.
constexpr uint32_t PERF_CELL_LEVELS = 4;                 // PeRF's grid: levels 0..3 are dense
constexpr uint64_t PERF_CELL_CAP = 1ull << 20;           // at most 2^20 cells (64 MB) are ever duplicated
struct PackedLayout {
    uint32_t n_cell_levels;                              // leading dense levels with a cell-major copy (<= PERF_CELL_LEVELS)
    uint64_t cell_start[PERF_CELL_LEVELS];
    uint64_t total_entries;                              // n_entries + 8 * cells
};
inline PackedLayout packed_layout(const LevelTable& lt, uint64_t n_entries)
{
    PackedLayout pl; pl.n_cell_levels = 0; pl.total_entries = n_entries;
    uint64_t cells = 0;
    for (uint32_t l = 0; l < lt.n_levels && l < PERF_CELL_LEVELS; ++l) {
        if ((lt.hashed_mask >> l) & 1u) break;
        const uint64_t c = (uint64_t)lt.res[l] * lt.res[l] * lt.res[l];
        if (cells + c > PERF_CELL_CAP) break;
        pl.cell_start[l] = pl.total_entries; pl.total_entries += 8 * c; cells += c; pl.n_cell_levels = l + 1;
    }
    return pl;
}
int mlp_param_count(const perf_mlp_cfg* mlp, uint64_t* count);
int check_mlp(const perf_mlp_cfg* mlp);

#ifdef __CUDACC__
// ---------------------------------------------------------------- device: hash-grid addressing
// tcnn pos_fract / grid_index / coherent-prime hash; mirrored 1:1 by
// oracle/hashgrid.py so that indices are bit-identical and the fp32 blend matches to the ulp.
struct Corner8 {
    uint32_t idx[8];     // absolute entry index (level offset included)
    float    w[8];       // trilinear weight, corner c: bit0=x, bit1=y, bit2=z
};

__host__ __device__ __forceinline__ uint32_t level_index(uint32_t gx, uint32_t gy, uint32_t gz,
                                                bool hashed, bool pow2, uint32_t res, uint32_t size)
{
    uint32_t idx;
    if (hashed) {
        idx = gx ^ (gy * 2654435761u) ^ (gz * 805459861u);
        idx = pow2 ? (idx & (size - 1u)) : (idx % size);
    } else {
        // dense stride walk; for these levels res^3 (rounded up to 8) == size so all three
        // dims participate.  (tcnn stops early only when stride > size, which implies hashed.)
        idx = gx + gy * res + gz * res * res;
        if (idx >= size) idx %= size;
    }
    return idx;
}

// round-to-nearest multiply that the compiler may not contract into an fma (device); plain IEEE multiply
// when the same source is compiled for the host by tests/host_harness.py
#ifdef __CUDA_ARCH__
#define PERF_FMUL_RN(a, b) __fmul_rn((a), (b))
#define PERF_FADD_RN(a, b) __fadd_rn((a), (b))
#define PERF_FSUB_RN(a, b) __fsub_rn((a), (b))
#define PERF_FDIV_RN(a, b) __fdiv_rn((a), (b))
#else
#define PERF_FMUL_RN(a, b) ((a) * (b))
#define PERF_FADD_RN(a, b) ((a) + (b))
#define PERF_FSUB_RN(a, b) ((a) - (b))
#define PERF_FDIV_RN(a, b) ((a) / (b))
#endif

__host__ __device__ __forceinline__ void level_corners(const LevelTable& lt, int l, float x, float y, float z, Corner8& c)
{
    const float scale = lt.scale[l];
    const uint32_t res = lt.res[l], size = lt.size[l], off = lt.offset[l];
    const bool hashed = (lt.hashed_mask >> l) & 1u, pow2 = (lt.pow2_mask >> l) & 1u;
    float px = fmaf(scale, x, 0.5f), py = fmaf(scale, y, 0.5f), pz = fmaf(scale, z, 0.5f);
    float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    uint32_t gx = (uint32_t)(int)fx, gy = (uint32_t)(int)fy, gz = (uint32_t)(int)fz;
    float wx = px - fx, wy = py - fy, wz = pz - fz;
    if (lt.smoothstep) {
        wx = wx * wx * (3.0f - 2.0f * wx); wy = wy * wy * (3.0f - 2.0f * wy); wz = wz * wz * (3.0f - 2.0f * wz);
    }
    const float ox = 1.0f - wx, oy = 1.0f - wy, oz = 1.0f - wz;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // weight = ((1 * ax) * ay) * az in this order (oracle/hashgrid.py::_corner_weights_indices)
        float w = PERF_FMUL_RN(PERF_FMUL_RN((k & 1) ? wx : ox, (k & 2) ? wy : oy), (k & 4) ? wz : oz);
        c.w[k] = w;
        c.idx[k] = off + level_index(gx + (k & 1), gy + ((k >> 1) & 1), gz + ((k >> 2) & 1), hashed, pow2, res, size);
    }
}

// Specialised addressing for the hot kernels: the level kind is a compile-time constant
// (HASHED levels must have a power-of-two size, dense levels use idx < 2*size), coordinates must
// lie in [0,1] (callers clamp masked-out samples) and interpolation is Linear.  Produces exactly
// the indices / weights of level_corners() under those preconditions -- no divisions, no branches.
template <bool HASHED>
__host__ __device__ __forceinline__ void level_corners_fast(const LevelTable& lt, int l, float x, float y, float z, Corner8& c)
{
    const float scale = lt.scale[l];
    const uint32_t res = lt.res[l], size = lt.size[l], off = lt.offset[l];
    const float px = fmaf(scale, x, 0.5f), py = fmaf(scale, y, 0.5f), pz = fmaf(scale, z, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const uint32_t gx = (uint32_t)(int)fx, gy = (uint32_t)(int)fy, gz = (uint32_t)(int)fz;
    const float wx = px - fx, wy = py - fy, wz = pz - fz;
    const float ox = 1.0f - wx, oy = 1.0f - wy, oz = 1.0f - wz;
    const float wxy[4] = {PERF_FMUL_RN(ox, oy), PERF_FMUL_RN(wx, oy), PERF_FMUL_RN(ox, wy), PERF_FMUL_RN(wx, wy)};
#pragma unroll
    for (int k = 0; k < 8; ++k) c.w[k] = PERF_FMUL_RN(wxy[k & 3], (k & 4) ? wz : oz);
    if constexpr (HASHED) {
        const uint32_t mask = size - 1u;
        const uint32_t hy0 = gy * 2654435761u, hy1 = hy0 + 2654435761u;
        const uint32_t hz0 = gz * 805459861u,  hz1 = hz0 + 805459861u;
        const uint32_t hyz[4] = {hy0 ^ hz0, hy1 ^ hz0, hy0 ^ hz1, hy1 ^ hz1};
        const uint32_t gx1 = gx + 1u;
#pragma unroll
        for (int k = 0; k < 8; ++k) c.idx[k] = off + ((((k & 1) ? gx1 : gx) ^ hyz[k >> 1]) & mask);
    } else {
        const uint32_t r2 = res * res;
        const uint32_t b00 = gx + gy * res + gz * r2;
        const uint32_t base[4] = {b00, b00 + res, b00 + r2, b00 + res + r2};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t i = base[k >> 1] + (uint32_t)(k & 1);
            i = (i >= size) ? i - size : i;                 // == i % size for i < 2*size
            c.idx[k] = off + i;
        }
    }
}
// Opaque bit operations (device: inline PTX the optimiser cannot re-associate; host harness: plain C).
// (a ^ b ^ c) & m is what the compiler makes of the hashed index -- per corner one 3-input XOR, one AND and one ADD of
// the level offset.  Masking the x term and the four (y ^ z) terms ONCE per level leaves a single 2-input XOR per corner.
__host__ __device__ __forceinline__ uint32_t xor_and_u32(uint32_t a, uint32_t b, uint32_t m)
{
#ifdef __CUDA_ARCH__
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(r) : "r"(a), "r"(b), "r"(m)); return r;     // (a ^ b) & m
#else
    return (a ^ b) & m;
#endif
}
__host__ __device__ __forceinline__ uint32_t and_u32(uint32_t a, uint32_t m)
{
#ifdef __CUDA_ARCH__
    uint32_t r; asm("and.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(m)); return r;
#else
    return a & m;
#endif
}

// level_corners_fast() with indices RELATIVE to the level's first entry: the caller adds lt.offset[l] to the table
// pointer once per level (idx_fast[k] == lt.offset[l] + idx[k], weights identical; tests/test_addressing_host.py).
// Hashed levels: 4 + 2 masked terms and one XOR per corner (see xor_and_u32).  Dense levels: the `% size` wrap can
// only trigger in cells on the far faces of the box, so ONE comparison per level guards the eight per-corner wraps.
template <bool HASHED>
__host__ __device__ __forceinline__ void level_corners_rel(const LevelTable& lt, int l, float x, float y, float z, uint32_t (&idx)[8], float (&w)[8])
{
    const float scale = lt.scale[l];
    const uint32_t res = lt.res[l], size = lt.size[l];
    const float px = fmaf(scale, x, 0.5f), py = fmaf(scale, y, 0.5f), pz = fmaf(scale, z, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const uint32_t gx = (uint32_t)(int)fx, gy = (uint32_t)(int)fy, gz = (uint32_t)(int)fz;
    const float wx = px - fx, wy = py - fy, wz = pz - fz;
    const float ox = 1.0f - wx, oy = 1.0f - wy, oz = 1.0f - wz;
    const float wxy[4] = {PERF_FMUL_RN(ox, oy), PERF_FMUL_RN(wx, oy), PERF_FMUL_RN(ox, wy), PERF_FMUL_RN(wx, wy)};
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = PERF_FMUL_RN(wxy[k & 3], (k & 4) ? wz : oz);
    if constexpr (HASHED) {
        const uint32_t mask = size - 1u;
        const uint32_t hy0 = gy * 2654435761u, hy1 = hy0 + 2654435761u;
        const uint32_t hz0 = gz * 805459861u,  hz1 = hz0 + 805459861u;
        const uint32_t hyz[4] = {xor_and_u32(hy0, hz0, mask), xor_and_u32(hy1, hz0, mask), xor_and_u32(hy0, hz1, mask), xor_and_u32(hy1, hz1, mask)};
        const uint32_t hx[2] = {and_u32(gx, mask), and_u32(gx + 1u, mask)};
#pragma unroll
        for (int k = 0; k < 8; ++k) idx[k] = hx[k & 1] ^ hyz[k >> 1];
    } else {
        const uint32_t r2 = res * res;
        const uint32_t b00 = gx + gy * res + gz * r2;
        const uint32_t base[4] = {b00, b00 + res, b00 + r2, b00 + res + r2};
#pragma unroll
        for (int k = 0; k < 8; ++k) idx[k] = base[k >> 1] + (uint32_t)(k & 1);
        if (base[3] + 1u >= size) {                             // the largest of the eight; every index is < 2*size
#pragma unroll
            for (int k = 0; k < 8; ++k) idx[k] = (idx[k] >= size) ? idx[k] - size : idx[k];
        }
    }
}
// Dense level of the fused field kernels: cell index into the cell-major copy (PackedLayout) + the eight weights of
// level_corners_fast().  Coordinates in [0,1] => gx, gy, gz <= res - 1, i.e. cell < res^3.
__host__ __device__ __forceinline__ uint32_t level_cell_dense(const LevelTable& lt, int l, float x, float y, float z, float (&w)[8])
{
    const float scale = lt.scale[l];
    const uint32_t res = lt.res[l];
    const float px = fmaf(scale, x, 0.5f), py = fmaf(scale, y, 0.5f), pz = fmaf(scale, z, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const uint32_t gx = (uint32_t)(int)fx, gy = (uint32_t)(int)fy, gz = (uint32_t)(int)fz;
    const float wx = px - fx, wy = py - fy, wz = pz - fz;
    const float ox = 1.0f - wx, oy = 1.0f - wy, oz = 1.0f - wz;
    const float wxy[4] = {PERF_FMUL_RN(ox, oy), PERF_FMUL_RN(wx, oy), PERF_FMUL_RN(ox, wy), PERF_FMUL_RN(wx, wy)};
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = PERF_FMUL_RN(wxy[k & 3], (k & 4) ? wz : oz);
    return gx + res * (gy + res * gz);
}

// host-side precondition of the fast path: `n_dense` leading dense levels, every other level
// hashed with a power-of-two size, Linear interpolation
inline bool fast_addressing_ok(const LevelTable& lt, uint32_t n_dense)
{
    if (lt.smoothstep) return false;
    for (uint32_t l = 0; l < lt.n_levels; ++l) {
        const bool hashed = (lt.hashed_mask >> l) & 1u, pow2 = (lt.pow2_mask >> l) & 1u;
        if (l < n_dense ? hashed : !(hashed && pow2)) return false;
    }
    return true;
}

// Scatter of one level's 8 corner contributions (float2 each) into an fp32 gradient table.
// V4: the two x-neighbours of a corner pair sit in the same 16-byte slot whenever their indices differ in bit 0
// only (always for a hashed level when the cell's x coordinate is even, and for a dense level when the index
// is even): one 16-byte vector atomic then replaces two 8-byte ones.  `dtable` must be 16-byte aligned for V4.
// (Host compilation = tests/host_harness.py: plain adds, one thread.)
__host__ __device__ __forceinline__ void grad_add2(float2* p, float2 v)
{
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    p->x += v.x; p->y += v.y;
#endif
}
__host__ __device__ __forceinline__ void grad_add4(float4* p, float4 v)
{
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w;
#endif
}
template <bool V4>
__host__ __device__ __forceinline__ void scatter8(float2* __restrict__ dtable, const uint32_t (&idx)[8], const float2 (&v)[8])
{
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const uint32_t i0 = idx[k], i1 = idx[k + 1];
        if (V4 && ((i0 ^ i1) == 1u)) {
            const bool swap = (i0 & 1u) != 0u;
            const float2 lo = swap ? v[k + 1] : v[k], hi = swap ? v[k] : v[k + 1];
            grad_add4(reinterpret_cast<float4*>(dtable + (i0 & ~1u)), make_float4(lo.x, lo.y, hi.x, hi.y));
        } else {
            grad_add2(dtable + i0, v[k]);
            grad_add2(dtable + i1, v[k + 1]);
        }
    }
}

__host__ __device__ __forceinline__ uint32_t pack_half2(float a, float b)
{
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__host__ __device__ __forceinline__ float2 unpack_half2(uint32_t u)
{
    return __half22float2(*reinterpret_cast<__half2*>(&u));
}
__device__ __forceinline__ float round_half(float v) { return __half2float(__float2half_rn(v)); }

// Trilinear blend of one level, exactly as tcnn's kernel_grid does it for __half parameters:
// result = fma((half)weight_k, value_k, result) for corners k = 0..7, one fp16 rounding per fma
// (HFMA2 on the feature pair), starting from zero.  Mirrored by oracle/hashgrid.py::encode(blend="half").
__device__ __forceinline__ uint32_t blend8_half(const float (&w)[8], const uint32_t (&v)[8])
{
    __half2 acc = __float2half2_rn(0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = __hfma2(__float2half2_rn(w[k]), *reinterpret_cast<const __half2*>(&v[k]), acc);
    return *reinterpret_cast<uint32_t*>(&acc);
}

// (p - lo) / ext, correctly rounded, for a divisor that is the same for every sample of the launch: with r = RN(1 / ext)
// computed once, q = RN(n r), q' = RN(q + RN(n - q ext) r) is the correctly rounded quotient (Markstein: the remainder is
// exact in one FMA; holds unless the significand of ext is all ones, which the host checks with div_uniform_ok) -- 3
// dependent FMA-pipe instructions per coordinate instead of the ~10 of the generic IEEE division (MUFU.RCP, refinement,
// range check).  __host__ __device__: tests/test_addressing_host.py compares the sequence with the IEEE division on the CPU.
#ifndef PERF_OPT_DIV
#define PERF_OPT_DIV 1
#endif
__host__ __device__ __forceinline__ float div_uniform(float n, float ext, float r, bool generic)
{
#if PERF_OPT_DIV
    if (generic) return PERF_FDIV_RN(n, ext);             // uniform branch; never taken for PeRF's [-1,1]^3 box
    const float q = PERF_FMUL_RN(n, r);
    return fmaf(fmaf(-q, ext, n), r, q);
#else
    (void)r; (void)generic; return PERF_FDIV_RN(n, ext);
#endif
}
// host-side precondition of the 3-FMA path: a normal-range extent whose significand is not all ones
inline bool div_uniform_ok(float ext)
{
    uint32_t bits; memcpy(&bits, &ext, 4);
    return (bits & 0x7FFFFFu) != 0x7FFFFFu && ext > 1e-30f && ext < 1e30f;
}

// ---------------------------------------------------------------- device: tcgen05 / mbarrier PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Spin on an mbarrier phase with a bounded number of polls; a kernel that would hang traps
// instead (so a descriptor bug surfaces as a CUDA error, not a wedged GPU).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if (spin == 64) t0 = clock64();
        if (spin > 64 && clock64() - t0 > 4000000000ll) __trap();    // ~2 s at 2 GHz: a lost MMA completion
    }
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst)   // one full warp
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(smem_dst)), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr)     // same warp that allocated
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(COLS) : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"): 8x16-byte core
// matrices; LBO = byte distance between the two core matrices of one K=16 step,
// SBO = byte distance between 8-row groups.  (cute::UMMA::SmemDescriptor, version_=1.)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    return d;                                     // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}
// Instruction descriptor: fp16 x fp16 -> fp32, both operands K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N)
{
    return (1u << 4)                 // D format  = F32
         | (0u << 7) | (0u << 10)    // A, B format = F16
         | (0u << 15) | (0u << 16)   // A, B K-major
         | ((uint32_t)(N >> 3) << 17)
         | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of warp w reads TMEM lane 32*(w%4)+i.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
#endif  // __CUDACC__

}  // namespace perf
