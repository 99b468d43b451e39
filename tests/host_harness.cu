// TEST HARNESS (compiled only by tests/host_harness.py into tests/_build/, never into libperfb200.so):
// host entry points that run the __host__ __device__ addressing helpers of perf_b200/csrc/common.cuh over
// host arrays, so the CPU test-suite can compare the generic and the specialised ("fast") hash-grid
// addressing with each other and with the oracle for arbitrary grid configurations.
#include "../perf_b200/csrc/common.cuh"

using namespace perf;

extern "C" {
#pragma GCC visibility push(default)

// mode 0: level_corners (generic); mode 1: level_corners_fast<HASHED> with HASHED taken from the level table.
// Returns 1 in *fast_ok when fast_addressing_ok(lt, n_dense) holds.  idx [N,8] uint32, w [N,8] float.
int perf_host_level_corners(const perf_grid_cfg* cfg, int level, int mode, uint32_t n_dense, const float* x01, uint64_t N,
                            uint32_t* idx, float* w, int* fast_ok)
{
    LevelTable lt; int rc = build_level_table(cfg, &lt, nullptr); if (rc) return rc;
    if (fast_ok) *fast_ok = fast_addressing_ok(lt, n_dense) ? 1 : 0;
    if (level < 0 || level >= (int)lt.n_levels) return PERF_EINVAL;
    const bool hashed = (lt.hashed_mask >> level) & 1u;
    for (uint64_t i = 0; i < N; ++i) {
        Corner8 c;
        if (mode == 0) level_corners(lt, level, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], c);
        else if (mode == 3) {                      // dense level of the fused field kernels: cell index (returned in idx[0]) + weights
            if (hashed) return PERF_EINVAL;
            const uint32_t cell = level_cell_dense(lt, level, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], c.w);
            for (int k = 0; k < 8; ++k) c.idx[k] = cell;
        }
        else if (mode == 2) {                      // level-local variant of the fused field kernels: offset added back here
            uint32_t ri[8];
            if (hashed) level_corners_rel<true>(lt, level, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], ri, c.w);
            else level_corners_rel<false>(lt, level, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], ri, c.w);
            for (int k = 0; k < 8; ++k) c.idx[k] = lt.offset[level] + ri[k];
        }
        else if (hashed) level_corners_fast<true>(lt, level, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], c);
        else level_corners_fast<false>(lt, level, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], c);
        for (int k = 0; k < 8; ++k) { idx[8 * i + k] = c.idx[k]; w[8 * i + k] = c.w[k]; }
    }
    return PERF_OK;
}

// div_uniform (common.cuh): out[i] = n[i] / ext by the 3-FMA sequence (or the IEEE division when the extent is not admitted);
// *ok = div_uniform_ok(ext)
int perf_host_div_uniform(const float* n, uint64_t N, float ext, float* out, int* ok)
{
    const bool admitted = div_uniform_ok(ext);
    if (ok) *ok = admitted ? 1 : 0;
    const float r = 1.0f / ext;                            // the correctly rounded reciprocal (device: __frcp_rn)
    for (uint64_t i = 0; i < N; ++i) out[i] = div_uniform(n[i], ext, r, !admitted);
    return PERF_OK;
}

// scatter8<V4> over N rows of (idx[8], v[8] float2) into dtable (float2 entries, 16-byte aligned for v4 != 0)
int perf_host_scatter8(int v4, const uint32_t* idx, const float* v, uint64_t N, float* dtable)
{
    if (v4 && ((uintptr_t)dtable % 16) != 0) return PERF_EINVAL;
    for (uint64_t i = 0; i < N; ++i) {
        uint32_t id[8]; float2 val[8];
        for (int k = 0; k < 8; ++k) { id[k] = idx[8 * i + k]; val[k] = make_float2(v[16 * i + 2 * k], v[16 * i + 2 * k + 1]); }
        if (v4) scatter8<true>((float2*)dtable, id, val); else scatter8<false>((float2*)dtable, id, val);
    }
    return PERF_OK;
}

#pragma GCC visibility pop
}
