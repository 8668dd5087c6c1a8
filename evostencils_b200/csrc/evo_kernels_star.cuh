// evo_kernels_star.cuh -- specialised high-bandwidth kernels for scalar real star stencils
// (7-point 3-D; 5-point 2-D).  Every try_* function returns false when the statement does not match
// its fast path; the caller then launches the generic kernel.  Results are bit-identical to the
// generic kernels and to the oracle (same operation order per node, -fmad=false).
//
// 3-D design (the HBM-bound regime, 513^3: 1.1 GB per field):
//   * one CTA owns an XY tile and marches through z; planes (tile + halo) are staged into a shared
//     memory ring by TMA (cp.async.bulk.tensor.3d + mbarrier), one plane of prefetch ahead
//   * red-black Gauss-Seidel: the 2*k half-sweeps of k consecutive sweeps are pipeline stages that lag
//     one plane each and shrink their XY/Z halo by one node per stage (redundant halo updates are
//     recomputed identically by the neighbouring CTA, so the result is the exact sequential RB-GS);
//     u is read once and written once per k sweeps: 24 B/DOF per launch instead of 48 per sweep
//   * the kernel is out of place (reads SOL, writes the [next] slot): another CTA may still need the
//     old values of a plane this CTA has already finished
#pragma once
#include <cuda.h>

#include <algorithm>
#include <type_traits>

#include "evo_kernels.cuh"

namespace evo {
namespace star {

// ---------------------------------------------------------------------------------------------
struct Star7 {  // coefficients in ascending table order: z-1, y-1, x-1, centre, x+1, y+1, z+1
    double zm, ym, xm, c, xp, yp, zp;
};

static bool match_star7(const Sten &s, Star7 *out)
{
    static const signed char ex[7][3] = {{0, 0, -1}, {0, -1, 0}, {-1, 0, 0}, {0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (s.nnz != 7) return false;
    for (int q = 0; q < 7; ++q)
        if (s.ox[q] != ex[q][0] || s.oy[q] != ex[q][1] || s.oz[q] != ex[q][2] || s.im[q] != 0.0) return false;
    out->zm = s.re[0]; out->ym = s.re[1]; out->xm = s.re[2]; out->c = s.re[3];
    out->xp = s.re[4]; out->yp = s.re[5]; out->zp = s.re[6];
    return true;
}

// ---------------------------------------------------------------------------------------------
// TMA plumbing (raw PTX; no CUTLASS)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled()
{
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 3-D fp64 tensor map over one padded field: dims (n, n, n), strides (pitch, plane), box (bx, by, 1);
// out-of-range box elements (negative coordinates, x >= n, ...) are zero filled
static bool make_plane_map(CUtensorMap *map, const Geom &g, const double *base, int box_x, int box_y)
{
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)g.n, (cuuint64_t)g.n, (cuuint64_t)g.nz};
    cuuint64_t strides[2] = {(cuuint64_t)g.pitch * 8, (cuuint64_t)g.plane * 8};
    cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_plane(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"((unsigned long long)map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// k consecutive pointwise RB-GS sweeps (S = 2k half-sweeps) in one pass, 3-D 7-point.
constexpr int RB_TX = 64, RB_TY = 16, RB_NT = 256;

template <int S>
__global__ void __launch_bounds__(RB_NT) k3_rbgs_stream(const __grid_constant__ CUtensorMap umap,
                                                        const double *__restrict__ f, double *__restrict__ uout,
                                                        const Geom g, const Star7 c, const double omega, const int tz)
{
    constexpr int H = S;                 // load halo
    // TMA needs a 16-byte aligned start address: with fp64 the x start coordinate must be even, so the
    // window starts one node further left (tiles start at odd x = 1 + 64*bx) and is 2 nodes wider
    constexpr int LX = RB_TX + 2 * H + 2, LY = RB_TY + 2 * H;
    constexpr int NP = S + 3;            // ring: planes t-S .. t+2 (one plane of TMA prefetch)
    constexpr uint32_t PLANE_BYTES = LX * LY * 8;
    constexpr int PSTRIDE = (LX * LY + 15) / 16 * 16;  // slot stride in doubles (128-byte aligned slots)
    extern __shared__ __align__(128) double ring[];
    __shared__ __align__(8) uint64_t bars[NP];

    const int n = g.n;
    const int x0 = 1 + blockIdx.x * RB_TX, y0 = 1 + blockIdx.y * RB_TY;
    const int za = 1 + blockIdx.z * tz, zb = min(za + tz - 1, n - 2);
    const int xb = x0 - H - 1, yb = y0 - H;  // global coordinates of local (0, 0); xb is even
    const int pbase = za - H;            // first plane ever loaded (may be < 0: zero filled, never used)
    const int tid = threadIdx.x;

    if (tid == 0) {
        for (int i = 0; i < NP; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto slot_of = [&](int p) { return (p - pbase) % NP; };
    auto issue = [&](int p) {  // one thread
        const int sl = slot_of(p);
        mbar_expect_tx(&bars[sl], PLANE_BYTES);
        tma_load_plane(ring + (size_t)sl * PSTRIDE, &umap, xb, yb, p, &bars[sl]);
    };
    auto wait_plane = [&](int p) { mbar_wait(&bars[slot_of(p)], (uint32_t)(((p - pbase) / NP) & 1)); };

    const int t0 = za - (S - 1);         // first plane of stage 0
    const int t1 = zb + (S - 1);         // last step: stage S-1 reaches plane zb
    const int pmax = min(zb + H, n - 1); // last plane that is ever read
    if (tid == 0) {
        for (int p = pbase; p <= min(t0 + 1, pmax); ++p) issue(p);
    }
    for (int p = pbase; p <= min(t0, pmax); ++p) wait_plane(p);

    for (int t = t0; t <= t1; ++t) {
        if (t + 1 <= pmax) wait_plane(t + 1);
        if (tid == 0 && t + 2 <= pmax) {
            fence_proxy_async();         // generic-proxy accesses of the recycled slot are complete (barrier below)
            issue(t + 2);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int z = t - s;
            const int e = S - 1 - s;     // halo this stage still has to cover
            if (z >= max(za - e, 1) && z <= min(zb + e, n - 2)) {
                const int xlo = max(x0 - e, 1), xhi = min(x0 + RB_TX - 1 + e, n - 2);
                const int ylo = max(y0 - e, 1), yhi = min(y0 + RB_TY - 1 + e, n - 2);
                const int color = s & 1;
                const int npx = ((xhi - xlo) >> 1) + 1, nrow = yhi - ylo + 1;
                double *pc = ring + (size_t)slot_of(z) * PSTRIDE;
                const double *pm = ring + (size_t)slot_of(z - 1) * PSTRIDE;
                const double *pp = ring + (size_t)slot_of(z + 1) * PSTRIDE;
                for (int i = tid; i < npx * nrow; i += RB_NT) {
                    const int ry = i / npx, k = i - ry * npx;
                    const int y = ylo + ry;
                    const int x = xlo + ((xlo + y + z + color) & 1) + 2 * k;
                    if (x > xhi) continue;
                    const int li = (y - yb) * LX + (x - xb);
                    const double fv = __ldg(f + node_index(g, x, y, z));
                    double sum = 0.0;
                    sum = sum + c.zm * pm[li];
                    sum = sum + c.ym * pc[li - LX];
                    sum = sum + c.xm * pc[li - 1];
                    sum = sum + c.xp * pc[li + 1];
                    sum = sum + c.yp * pc[li + LX];
                    sum = sum + c.zp * pp[li];
                    const double xs = (fv - sum) / c.c;
                    const double old = pc[li];
                    pc[li] = old + omega * (xs - old);
                }
            }
            __syncthreads();
        }
        const int zf = t - (S - 1);
        if (zf >= za && zf <= zb) {
            const double *pc = ring + (size_t)slot_of(zf) * PSTRIDE;
            const int xhi = min(x0 + RB_TX - 1, n - 2), yhi = min(y0 + RB_TY - 1, n - 2);
            for (int i = tid; i < RB_TX * RB_TY; i += RB_NT) {
                const int ry = i / RB_TX, rx = i - ry * RB_TX;
                const int x = x0 + rx, y = y0 + ry;
                if (x <= xhi && y <= yhi) uout[node_index(g, x, y, zf)] = pc[(y - yb) * LX + (x - xb)];
            }
        }
        // no barrier needed here: the slot the next step's TMA recycles (plane t-S) was last read by stage
        // S-1 above, i.e. before that stage's barrier; the store loop reads a different slot
    }
}

template <int S>
static bool launch_rbgs_stream(int sm_count, const Geom &g, const Star7 &c, const double *u, const double *f, double *uout,
                               double omega, cudaStream_t s)
{
    constexpr int H = S, LX = RB_TX + 2 * H + 2, LY = RB_TY + 2 * H, NP = S + 3;
    constexpr int PSTRIDE = (LX * LY + 15) / 16 * 16;
    CUtensorMap map;
    if (!make_plane_map(&map, g, u, LX, LY)) return false;
    const size_t smem = (size_t)NP * PSTRIDE * 8;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k3_rbgs_stream<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        attr_set = true;
    }
    const int inner = g.n - 2;
    const int tx = (inner + RB_TX - 1) / RB_TX, ty = (inner + RB_TY - 1) / RB_TY;
    // z slabs: enough CTAs for ~4 per SM, but slabs of at least 32 planes (pipeline fill = S-1 planes)
    int slabs = std::max(1, std::min(inner / 32, (4 * sm_count + tx * ty - 1) / (tx * ty)));
    const int tz = (inner + slabs - 1) / slabs;
    slabs = (inner + tz - 1) / tz;
    k3_rbgs_stream<S><<<dim3(tx, ty, slabs), RB_NT, smem, s>>>(map, f, uout, g, c, omega, tz);
    return cudaGetLastError() == cudaSuccess;
}

// k pointwise RB-GS sweeps, out of place (u -> uout); returns false if not applicable
template <typename T, int DIM, int NF>
static bool try_rbgs_stream(int sm_count, const Geom &g, const OpSten &st, Fields<T> u, Fields<T> f, Fields<T> uout,
                            double omega, int sweeps, cudaStream_t s)
{
    if constexpr (std::is_same<T, double>::value && DIM == 3 && NF == 1) {
        Star7 c;
        if (g.n < 33 || !match_star7(st.s[0][0], &c)) return false;
        if (sweeps == 1) return launch_rbgs_stream<2>(sm_count, g, c, u.p[0], f.p[0], uout.p[0], omega, s);
        if (sweeps == 2) return launch_rbgs_stream<4>(sm_count, g, c, u.p[0], f.p[0], uout.p[0], omega, s);
        return false;
    } else {
        return false;
    }
}

template <typename T, int DIM, int NF>
static bool rbgs_stream_applicable(const Geom &g, const OpSten &st)
{
    if constexpr (std::is_same<T, double>::value && DIM == 3 && NF == 1) {
        Star7 c;
        return g.n >= 33 && match_star7(st.s[0][0], &c) && get_encode_tiled() != nullptr;
    } else {
        return false;
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T, int DIM, int NF>
static bool try_residual(int, const Geom &, const OpSten &, Fields<T>, Fields<T>, Fields<T>, cudaStream_t)
{
    return false;
}
template <typename T, int DIM, int NF>
static bool try_smooth_point(int, const Geom &, const OpSten &, const SmoothParams &, Fields<T>, Fields<T>, Fields<T>,
                             cudaStream_t)
{
    return false;
}
template <typename T, int DIM, int NF>
static bool try_residual_restrict(int, const Geom &, const Geom &, const OpSten &, const TransferW &, Fields<T>, Fields<T>,
                                  Fields<T>, cudaStream_t)
{
    return false;
}
template <typename T, int DIM, int NF>
static bool try_prolong_add(int, const Geom &, const Geom &, const TransferW &, Fields<T>, Fields<T>, double, cudaStream_t)
{
    return false;
}

}  // namespace star
}  // namespace evo
