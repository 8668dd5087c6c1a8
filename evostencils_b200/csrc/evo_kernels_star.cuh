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
#include <cstdlib>
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
constexpr int RB_TX = 64;  // tile width (x); tile height TY and thread count NT are template parameters

// compile-time geometry of pipeline stage s of S: it covers the tile plus a halo of E nodes
template <int S, int TY, int NT, int s> struct StageCfg {
    static constexpr int E = S - 1 - s;
    static constexpr int W = RB_TX + 2 * E, ROWS = TY + 2 * E, NPX = W / 2;
    // Bank-conflict-free lane mapping: the active nodes of one row all have the same x parity and would
    // only hit every other 8-byte bank; a group of 16 lanes (one shared-memory wavefront of 64-bit
    // accesses) therefore takes 8 consecutive active nodes of row y and 8 of row y+1 -- the two rows'
    // active colours have opposite x parity and together cover all 16 banks.
    static constexpr int CH = (NPX + 7) / 8, ROWP = (ROWS + 1) / 2;
    static constexpr int ITEMS = ROWP * CH * 16, ROUNDS = (ITEMS + NT - 1) / NT;
};
template <int S, int TY, int NT, int s> __host__ __device__ constexpr int stage_base()
{
    if constexpr (s == 0) return 0;
    else return stage_base<S, TY, NT, s - 1>() + StageCfg<S, TY, NT, s - 1>::ROUNDS;
}
template <int S, int TY, int NT> struct RbCfg {
    static constexpr int H = S;
    // TMA needs a 16-byte aligned start address: with fp64 the x start coordinate must be even, so the
    // window starts one node further left (tiles start at odd x = 1 + 64*bx) and is 2 nodes wider
    static constexpr int LX = RB_TX + 2 * H + 2, LY = TY + 2 * H;
    static constexpr int NP = S + 3;  // ring: planes t-S .. t+2 (one plane of TMA prefetch)
    static constexpr int PSTRIDE = (LX * LY + 15) / 16 * 16;  // slot stride in doubles (128-byte aligned)
    static constexpr int TOT = stage_base<S, TY, NT, S>();             // work items per thread over all stages
};

// per-thread work items: fixed (row, x-pair) positions for every stage, set up once before the z loop
template <int S, int TY, int NT> struct RbItems {
    int loff[RbCfg<S, TY, NT>::TOT];   // shared-memory offset of the pair's left node
    int goff[RbCfg<S, TY, NT>::TOT];   // in-plane global offset of the pair's left node
    unsigned v0, v1, par;      // bit i: item i valid when the active node is the left / right one; parity bit
};

template <int S, int TY, int NT, int s>
__device__ __forceinline__ void rb_setup(RbItems<S, TY, NT> &it, int tid, int x0, int y0, int xb, int yb, const Geom &g)
{
    using C = StageCfg<S, TY, NT, s>;
    constexpr int base = stage_base<S, TY, NT, s>();
#pragma unroll
    for (int j = 0; j < C::ROUNDS; ++j) {
        const int i = tid + NT * j;
        const int grp = i >> 4, l16 = i & 15;
        const int rp = grp / C::CH, ch = grp - rp * C::CH;
        const int ry = 2 * rp + (l16 >> 3), kx = ch * 8 + (l16 & 7);
        const int y = y0 - C::E + ry, xw = x0 - C::E + 2 * kx;
        const bool ok = i < C::ITEMS && kx < C::NPX && ry < C::ROWS && y >= 1 && y <= g.n - 2;
        it.loff[base + j] = (y - yb) * RbCfg<S, TY, NT>::LX + (xw - xb);
        it.goff[base + j] = y * g.pitch + xw;
        if (ok && xw >= 1 && xw <= g.n - 2) it.v0 |= 1u << (base + j);
        if (ok && xw + 1 >= 1 && xw + 1 <= g.n - 2) it.v1 |= 1u << (base + j);
        if ((xw + y + (s & 1)) & 1) it.par |= 1u << (base + j);   // parity of the left node for z even
    }
    if constexpr (s + 1 < S) rb_setup<S, TY, NT, s + 1>(it, tid, x0, y0, xb, yb, g);
}

// f values of the active nodes of every stage for the step whose stage-0 plane is t
template <int S, int TY, int NT, int s>
__device__ __forceinline__ void rb_load_f(const RbItems<S, TY, NT> &it, double (&fv)[RbCfg<S, TY, NT>::TOT], const double *__restrict__ f,
                                          const Geom &g, int t, int za, int zb)
{
    using C = StageCfg<S, TY, NT, s>;
    constexpr int base = stage_base<S, TY, NT, s>();
    const int z = t - s;
    if (z >= max(za - C::E, g.zin0) && z <= min(zb + C::E, g.zin1)) {
        const double *fp = f + (long long)z * g.plane;
#pragma unroll
        for (int j = 0; j < C::ROUNDS; ++j) {
            const int idx = base + j;
            // active node = left node iff its colour (x+y+z+c) is even
            const unsigned p = ((it.par >> idx) ^ (unsigned)(z + g.zpar)) & 1u;
            const bool ok = ((p ? it.v1 : it.v0) >> idx) & 1u;
            if (ok) fv[idx] = __ldg(fp + it.goff[idx] + p);
        }
    }
    if constexpr (s + 1 < S) rb_load_f<S, TY, NT, s + 1>(it, fv, f, g, t, za, zb);
}

template <int S, int TY, int NT, int s>
__device__ __forceinline__ void rb_stages(const RbItems<S, TY, NT> &it, const double (&fv)[RbCfg<S, TY, NT>::TOT], double *ring, int pbase,
                                          const Geom &g, const Star7 &c, double inv_c, double omega, int t, int za, int zb)
{
    using C = StageCfg<S, TY, NT, s>;
    using R = RbCfg<S, TY, NT>;
    constexpr int base = stage_base<S, TY, NT, s>();
    const int z = t - s;
    if (z >= max(za - C::E, g.zin0) && z <= min(zb + C::E, g.zin1)) {
        double *pc = ring + (size_t)((z - pbase) % R::NP) * R::PSTRIDE;
        const double *pm = ring + (size_t)((z - 1 - pbase) % R::NP) * R::PSTRIDE;
        const double *pp = ring + (size_t)((z + 1 - pbase) % R::NP) * R::PSTRIDE;
#pragma unroll
        for (int j = 0; j < C::ROUNDS; ++j) {
            const int idx = base + j;
            const unsigned p = ((it.par >> idx) ^ (unsigned)(z + g.zpar)) & 1u;
            const bool ok = ((p ? it.v1 : it.v0) >> idx) & 1u;
            if (ok) {
                const int li = it.loff[idx] + (int)p;
                double sum = 0.0;
                sum = sum + c.zm * pm[li];
                sum = sum + c.ym * pc[li - R::LX];
                sum = sum + c.xm * pc[li - 1];
                sum = sum + c.xp * pc[li + 1];
                sum = sum + c.yp * pc[li + R::LX];
                sum = sum + c.zp * pp[li];
                const double xs = (fv[idx] - sum) * inv_c;
                const double old = pc[li];
                pc[li] = old + omega * (xs - old);
            }
        }
    }
    __syncthreads();
    if constexpr (s + 1 < S) rb_stages<S, TY, NT, s + 1>(it, fv, ring, pbase, g, c, inv_c, omega, t, za, zb);
}

template <int S, int TY, int NT>
__global__ void __launch_bounds__(NT) k3_rbgs_stream(const __grid_constant__ CUtensorMap umap,
                                                        const double *__restrict__ f, double *__restrict__ uout,
                                                        const Geom g, const Star7 c, const double inv_c,
                                                        const double omega, const int tz)
{
    using R = RbCfg<S, TY, NT>;
    constexpr int H = R::H, LX = R::LX, LY = R::LY, NP = R::NP, PSTRIDE = R::PSTRIDE, TOT = R::TOT;
    constexpr uint32_t PLANE_BYTES = LX * LY * 8;
    static_assert(TOT <= 32, "validity bit masks are 32 bits wide");
    extern __shared__ __align__(128) double ring[];
    __shared__ __align__(8) uint64_t bars[NP];

    const int n = g.n;
    const int x0 = 1 + blockIdx.x * RB_TX, y0 = 1 + blockIdx.y * TY;
    const int za = g.zlo + blockIdx.z * tz, zb = min(za + tz - 1, g.zhi);
    const int xb = x0 - H - 1, yb = y0 - H;  // global coordinates of local (0, 0); xb is even
    const int pbase = za - H;                // first plane ever loaded (may be < 0: zero filled, never used)
    const int tid = threadIdx.x;

    if (tid == 0) {
        for (int i = 0; i < NP; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    RbItems<S, TY, NT> it;
    it.v0 = it.v1 = it.par = 0u;
    rb_setup<S, TY, NT, 0>(it, tid, x0, y0, xb, yb, g);
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
    const int pmax = min(zb + H, g.nz - 1); // last plane that is ever read
    if (tid == 0) {
        for (int p = pbase; p <= min(t0 + 1, pmax); ++p) issue(p);
    }
    double fcur[TOT], fnext[TOT];
#pragma unroll
    for (int i = 0; i < TOT; ++i) { fcur[i] = 0.0; fnext[i] = 0.0; }
    rb_load_f<S, TY, NT, 0>(it, fcur, f, g, t0, za, zb);
    for (int p = pbase; p <= min(t0, pmax); ++p) wait_plane(p);

    for (int t = t0; t <= t1; ++t) {
        // right-hand sides of the next step: issued now, consumed one step later (latency hidden)
        if (t < t1) rb_load_f<S, TY, NT, 0>(it, fnext, f, g, t + 1, za, zb);
        if (t + 1 <= pmax) wait_plane(t + 1);
        if (tid == 0 && t + 2 <= pmax) {
            fence_proxy_async();         // generic-proxy accesses of the recycled slot completed before the last barrier
            issue(t + 2);
        }
        rb_stages<S, TY, NT, 0>(it, fcur, ring, pbase, g, c, inv_c, omega, t, za, zb);
        const int zf = t - (S - 1);
        if (zf >= za && zf <= zb) {
            const double *pc = ring + (size_t)slot_of(zf) * PSTRIDE;
            const int xhi = min(x0 + RB_TX - 1, n - 2), yhi = min(y0 + TY - 1, n - 2);
            double *op = uout + (long long)zf * g.plane;
#pragma unroll
            for (int i = tid; i < RB_TX * TY; i += NT) {
                const int ry = i / RB_TX, rx = i - ry * RB_TX;
                const int x = x0 + rx, y = y0 + ry;
                if (x <= xhi && y <= yhi) op[y * g.pitch + x] = pc[(y - yb) * LX + (x - xb)];
            }
        }
        // no barrier needed here: the slot the next step's TMA recycles (plane t-S) was last read by stage
        // S-1 above, i.e. before that stage's barrier; the store loop reads a different slot
#pragma unroll
        for (int i = 0; i < TOT; ++i) fcur[i] = fnext[i];
    }
}

// ---------------------------------------------------------------------------------------------
// Lean single-sweep variant of k3_rbgs_stream (S = 2): the same tiles, TMA ring, stages and lane mapping, but the
// z loop carries no index arithmetic -- per work item one 32-bit shared-memory byte offset (+8 when the right node
// of the pair is the active one), ring-slot bases advanced incrementally, ld/st.shared with immediate neighbour
// offsets, the copy-out offsets fixed per thread.  (The generic kernel spent ~80 % of its issue slots on integer
// and uniform-datapath instructions: ncu profiles/r1_c_*.)
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
template <int OFF> __device__ __forceinline__ double lds_f64_off(uint32_t addr)
{
    double v;
    if constexpr (OFF >= 0) asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
    else asm volatile("ld.shared.f64 %0, [%1+-%2];" : "=d"(v) : "r"(addr), "n"(-OFF));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

#ifndef RB_LEAN_MINB
#define RB_LEAN_MINB 6
#endif
constexpr int RB_LEAN_NP = 5;   // ring depth of the lean kernel: planes t-2 .. t+2 (6 = two planes of TMA prefetch: no gain measured)
template <int TY, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256 && TY <= 8) ? RB_LEAN_MINB : 1) k3_rbgs_lean(const __grid_constant__ CUtensorMap umap, const double *__restrict__ f,
                                                   double *__restrict__ uout, const Geom g, const Star7 c, const double inv_c,
                                                   const double omega, const int tz)
{
    using R = RbCfg<2, TY, NT>;
    using C0 = StageCfg<2, TY, NT, 0>;
    using C1 = StageCfg<2, TY, NT, 1>;
    constexpr int H = R::H, LX = R::LX, LY = R::LY, NP = RB_LEAN_NP, PSTRIDE = R::PSTRIDE;
    constexpr int PF = NP - 4;   // planes of TMA prefetch beyond t+1 (ring: planes t-2 .. t+1+PF)
    constexpr int R0 = C0::ROUNDS, R1 = C1::ROUNDS, TOT = R0 + R1;
    constexpr uint32_t PLANE_BYTES = LX * LY * 8, PB = PSTRIDE * 8;
    constexpr int OUTR = (RB_TX * TY + NT - 1) / NT;
    static_assert(TOT <= 32, "validity bit masks are 32 bits wide");
    extern __shared__ __align__(128) double ring[];
    __shared__ __align__(8) uint64_t bars[NP];

    const int n = g.n;
    const int x0 = 1 + blockIdx.x * RB_TX, y0 = 1 + blockIdx.y * TY;
    const int za = g.zlo + blockIdx.z * tz, zb = min(za + tz - 1, g.zhi);
    const int xb = x0 - H - 1, yb = y0 - H;  // global coordinates of local (0, 0); xb is even
    const int pbase = za - H;                // first plane ever loaded (may be < 0: zero filled, never used)
    const int tid = threadIdx.x;
    const uint32_t ring_u32 = smem_u32(ring);

    if (tid == 0) {
        for (int i = 0; i < NP; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    // work items (the mapping of rb_setup): byte offset of the pair's left node in a plane, its global in-plane
    // offset, validity of the left / right node, colour parity of the left node for even z
    uint32_t loff[TOT];
    int goff[TOT];
    unsigned v0 = 0u, v1 = 0u, par = 0u;
    auto setup = [&](auto cfg, int base, int sidx) {
        using C = decltype(cfg);
#pragma unroll
        for (int j = 0; j < C::ROUNDS; ++j) {
            const int i = tid + NT * j;
            const int grp = i >> 4, l16 = i & 15;
            const int rp = grp / C::CH, ch = grp - rp * C::CH;
            const int ry = 2 * rp + (l16 >> 3), kx = ch * 8 + (l16 & 7);
            const int y = y0 - C::E + ry, xw = x0 - C::E + 2 * kx;
            const bool ok = i < C::ITEMS && kx < C::NPX && ry < C::ROWS && y >= 1 && y <= n - 2;
            loff[base + j] = (uint32_t)(((y - yb) * LX + (xw - xb)) * 8);
            goff[base + j] = y * g.pitch + xw;
            if (ok && xw >= 1 && xw <= n - 2) v0 |= 1u << (base + j);
            if (ok && xw + 1 >= 1 && xw + 1 <= n - 2) v1 |= 1u << (base + j);
            if ((xw + y + sidx) & 1) par |= 1u << (base + j);
        }
    };
    setup(C0{}, 0, 0);
    setup(C1{}, R0, 1);
    // validity masks for q = (z + zpar) & 1 of the item's plane: active node = left iff its colour is even
    const unsigned okq0 = (~par & v0) | (par & v1), okq1 = (~par & v1) | (par & v0);
    // copy-out: fixed (row, x) elements of the tile per thread
    uint32_t co_s[OUTR];
    int co_g[OUTR];
#pragma unroll
    for (int k = 0; k < OUTR; ++k) {
        const int e = tid + k * NT;
        const int ry = e / RB_TX, rx = e - ry * RB_TX;
        const bool ok = e < RB_TX * TY && x0 + rx <= n - 2 && y0 + ry <= n - 2;
        co_s[k] = (uint32_t)(((ry + H) * LX + rx + H + 1) * 8);
        co_g[k] = ok ? (y0 + ry) * g.pitch + x0 + rx : -1;
    }
    __syncthreads();

    auto slot_of = [&](int p) { return (p - pbase) % NP; };
    auto issue = [&](int p) {  // one thread
        const int sl = slot_of(p);
        mbar_expect_tx(&bars[sl], PLANE_BYTES);
        tma_load_plane(ring + (size_t)sl * PSTRIDE, &umap, xb, yb, p, &bars[sl]);
    };
    const int t0 = za - 1, t1 = zb + 1;       // stage 0 works on plane t, stage 1 on plane t-1
    const int pmax = min(zb + H, g.nz - 1);   // last plane that is ever read
    if (tid == 0) {
        for (int p = pbase; p <= min(t0 + PF, pmax); ++p) issue(p);
    }
    const int lo0 = max(za - 1, g.zin0), hi0 = min(zb + 1, g.zin1);   // planes stage 0 updates (halo recomputation)
    // The z loop is unrolled by two (v = (t - t0) & 1): the colour parity of a plane alternates, so per v every
    // item has a loop-invariant byte offset of its active node, validity bit and right-hand-side offset.
    uint32_t off[2][TOT];
    int fsel[2][TOT];
    unsigned okv[2];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        const unsigned q0 = (unsigned)(t0 + v + g.zpar) & 1u;
        const unsigned pp0 = q0 ? ~par : par, pp1 = q0 ? par : ~par;     // stage 0 works on plane t, stage 1 on t-1
        const unsigned m0 = (1u << R0) - 1u;
        okv[v] = ((q0 ? okq1 : okq0) & m0) | ((q0 ? okq0 : okq1) & ~m0);
#pragma unroll
        for (int j = 0; j < TOT; ++j) {
            const unsigned p = ((j < R0 ? pp0 : pp1) >> j) & 1u;
            off[v][j] = loff[j] + (p << 3);
            fsel[v][j] = (int)p;
        }
    }
    // right-hand-side pointers of the items for the NEXT load (stage 0: plane t, stage 1: plane t-1), advanced per step
    const double *fptr[TOT];
#pragma unroll
    for (int j = 0; j < TOT; ++j) fptr[j] = f + (long long)(j < R0 ? t0 : t0 - 1) * g.plane + goff[j];
    double fbuf[2][TOT];      // right-hand sides of the current step (index v) and of the next one (1 - v)
#pragma unroll
    for (int i = 0; i < TOT; ++i) { fbuf[0][i] = 0.0; fbuf[1][i] = 0.0; }
    {   // right-hand sides of the first step
        const bool in0 = t0 >= lo0 && t0 <= hi0, in1 = t0 - 1 >= za && t0 - 1 <= zb;
#pragma unroll
        for (int j = 0; j < TOT; ++j) {
            if ((j < R0 ? in0 : in1) && ((okv[0] >> j) & 1u)) fbuf[0][j] = __ldg(fptr[j] + fsel[0][j]);
            fptr[j] += g.plane;
        }
    }
    int sl = slot_of(t0);                      // ring slot of plane t
    for (int p = pbase; p <= min(t0, pmax); ++p) mbar_wait(&bars[slot_of(p)], 0u);
    int wsl = slot_of(t0 + 1), wph = ((t0 + 1 - pbase) / NP) & 1;   // mbarrier slot / phase of plane t+1

    for (int tt = t0; tt <= t1; tt += 2) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int t = tt + v;
            if (t > t1) break;
            if (t < t1) {   // right-hand sides of the next step: issued now, consumed one step later (latency hidden)
                const bool in0 = t + 1 >= lo0 && t + 1 <= hi0, in1 = t >= za && t <= zb;
#pragma unroll
                for (int j = 0; j < TOT; ++j) {
                    if ((j < R0 ? in0 : in1) && ((okv[1 - v] >> j) & 1u)) fbuf[1 - v][j] = __ldg(fptr[j] + fsel[1 - v][j]);
                    fptr[j] += g.plane;
                }
            }
            if (t + 1 <= pmax) mbar_wait(&bars[wsl], (uint32_t)wph);
            if (tid == 0 && t + 1 + PF <= pmax) {
                fence_proxy_async();           // generic-proxy accesses of the recycled slot completed before the last barrier
                issue(t + 1 + PF);
            }
            const int s_m1 = sl == 0 ? NP - 1 : sl - 1, s_m2 = s_m1 == 0 ? NP - 1 : s_m1 - 1, s_p1 = sl + 1 == NP ? 0 : sl + 1;
            const uint32_t b_t = ring_u32 + sl * PB, b_m1 = ring_u32 + s_m1 * PB, b_m2 = ring_u32 + s_m2 * PB, b_p1 = ring_u32 + s_p1 * PB;
            if (t >= lo0 && t <= hi0) {        // stage 0: first colour on plane t
#pragma unroll
                for (int j = 0; j < R0; ++j)
                    if ((okv[v] >> j) & 1u) {
                        const uint32_t o = off[v][j];
                        const uint32_t ac = b_t + o;
                        double sum = 0.0;
                        sum = sum + c.zm * lds_f64(b_m1 + o);
                        sum = sum + c.ym * lds_f64_off<-LX * 8>(ac);
                        sum = sum + c.xm * lds_f64_off<-8>(ac);
                        sum = sum + c.xp * lds_f64_off<8>(ac);
                        sum = sum + c.yp * lds_f64_off<LX * 8>(ac);
                        sum = sum + c.zp * lds_f64(b_p1 + o);
                        const double xs = (fbuf[v][j] - sum) * inv_c;
                        const double old = lds_f64(ac);
                        sts_f64(ac, old + omega * (xs - old));
                    }
            }
            __syncthreads();
            if (t - 1 >= za && t - 1 <= zb) {  // stage 1: second colour on plane t-1, then the plane is final
#pragma unroll
                for (int j = R0; j < TOT; ++j)
                    if ((okv[v] >> j) & 1u) {
                        const uint32_t o = off[v][j];
                        const uint32_t ac = b_m1 + o;
                        double sum = 0.0;
                        sum = sum + c.zm * lds_f64(b_m2 + o);
                        sum = sum + c.ym * lds_f64_off<-LX * 8>(ac);
                        sum = sum + c.xm * lds_f64_off<-8>(ac);
                        sum = sum + c.xp * lds_f64_off<8>(ac);
                        sum = sum + c.yp * lds_f64_off<LX * 8>(ac);
                        sum = sum + c.zp * lds_f64(b_t + o);
                        const double xs = (fbuf[v][j] - sum) * inv_c;
                        const double old = lds_f64(ac);
                        sts_f64(ac, old + omega * (xs - old));
                    }
                __syncthreads();
                double *op = uout + (long long)(t - 1) * g.plane;
#pragma unroll
                for (int k = 0; k < OUTR; ++k)
                    if (co_g[k] >= 0) op[co_g[k]] = lds_f64(b_m1 + co_s[k]);
            } else {
                __syncthreads();
            }
            // no barrier needed here: the slot the next step's TMA recycles (plane t-2) was last read by stage 1
            // above, i.e. before that stage's barrier; the store loop reads a different slot
            sl = s_p1;
            if (++wsl == NP) { wsl = 0; wph ^= 1; }
        }
    }
}

template <int TY, int NT>
static bool launch_rbgs_lean(int sm_count, const Geom &g, const Star7 &c, const double *u, const double *f, double *uout,
                             double omega, cudaStream_t s)
{
    using R = RbCfg<2, TY, NT>;
    CUtensorMap map;
    if (!make_plane_map(&map, g, u, R::LX, R::LY)) return false;
    const size_t smem = (size_t)RB_LEAN_NP * R::PSTRIDE * 8;
    static int occ = 0;
    if (occ == 0) {
        if (cudaFuncSetAttribute(k3_rbgs_lean<TY, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3_rbgs_lean<TY, NT>, NT, smem) != cudaSuccess || occ < 1) occ = 1;
    }
    const int inner = g.n - 2, planes = g.zhi - g.zlo + 1;
    if (planes <= 0) return true;
    const int tx = (inner + RB_TX - 1) / RB_TX, ty = (inner + TY - 1) / TY;
    const long long slots = (long long)occ * sm_count;
    int best = 1;
    double best_cost = 1e300;
    for (int slabs = 1; slabs <= 16 && (slabs == 1 || slabs * 8 <= planes); ++slabs) {
        const int tzc = (planes + slabs - 1) / slabs;
        const long long ctas = (long long)tx * ty * ((planes + tzc - 1) / tzc);
        const double waves = (double)((ctas + slots - 1) / slots);
        const double cost = waves * (tzc + 6);
        if (cost < best_cost) { best_cost = cost; best = slabs; }
    }
    const int tz = (planes + best - 1) / best;
    const int slabs = (planes + tz - 1) / tz;
    k3_rbgs_lean<TY, NT><<<dim3(tx, ty, slabs), NT, smem, s>>>(map, f, uout, g, c, 1.0 / c.c, omega, tz);
    return cudaGetLastError() == cudaSuccess;
}

template <int S, int TY, int NT>
static bool launch_rbgs_stream(int sm_count, const Geom &g, const Star7 &c, const double *u, const double *f, double *uout,
                               double omega, cudaStream_t s)
{
    using R = RbCfg<S, TY, NT>;
    CUtensorMap map;
    if (!make_plane_map(&map, g, u, R::LX, R::LY)) return false;
    const size_t smem = (size_t)R::NP * R::PSTRIDE * 8;
    static int occ = 0;
    if (occ == 0) {
        if (cudaFuncSetAttribute(k3_rbgs_stream<S, TY, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3_rbgs_stream<S, TY, NT>, NT, smem) != cudaSuccess || occ < 1) occ = 1;
    }
    const int inner = g.n - 2, planes = g.zhi - g.zlo + 1;
    if (planes <= 0) return true;
    const int tx = (inner + RB_TX - 1) / RB_TX, ty = (inner + TY - 1) / TY;
    // z slabs: minimise (waves) x (planes per slab + pipeline fill) over 1..16 slabs
    const long long slots = (long long)occ * sm_count;
    int best = 1;
    double best_cost = 1e300;
    for (int slabs = 1; slabs <= 16 && (slabs == 1 || slabs * 8 <= planes); ++slabs) {
        const int tzc = (planes + slabs - 1) / slabs;
        const long long ctas = (long long)tx * ty * ((planes + tzc - 1) / tzc);
        const double waves = (double)((ctas + slots - 1) / slots);
        const double cost = waves * (tzc + 2 * S + 2);
        if (cost < best_cost) { best_cost = cost; best = slabs; }
    }
    const int tz = (planes + best - 1) / best;
    const int slabs = (planes + tz - 1) / tz;
    k3_rbgs_stream<S, TY, NT><<<dim3(tx, ty, slabs), NT, smem, s>>>(map, f, uout, g, c, 1.0 / c.c, omega, tz);
    return cudaGetLastError() == cudaSuccess;
}

// register-carried pair-column kernel (evo_kernels_rbcol.cuh)
template <int TY, int NPS, bool RC, int MINB>
static bool launch_rbgs_col(int sm_count, const Geom &g, const Star7 &c, const double *u, const double *f, double *uout, double omega,
                            cudaStream_t s);

// k pointwise RB-GS sweeps, out of place (u -> uout); returns false if not applicable
template <typename T, int DIM, int NF>
static bool try_rbgs_stream(int sm_count, const Geom &g, const OpSten &st, Fields<T> u, Fields<T> f, Fields<T> uout,
                            double omega, int sweeps, cudaStream_t s)
{
    if constexpr (std::is_same<T, double>::value && DIM == 3 && NF == 1) {
        Star7 c;
        if (g.n < 33 || !match_star7(st.s[0][0], &c)) return false;
        const int variant = option(OPT_RB_VARIANT);   // tile-shape / kernel experiments (0 = default)
        const double *up = u.p[0], *fp = f.p[0];
        double *op = uout.p[0];
        if (sweeps == 1) {
            if (variant == 1) return launch_rbgs_stream<2, 32, 512>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 2) return launch_rbgs_stream<2, 32, 256>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 3) return launch_rbgs_stream<2, 16, 256>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 4) return launch_rbgs_stream<2, 8, 128>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 5) return launch_rbgs_stream<2, 12, 256>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 6) return launch_rbgs_stream<2, 6, 128>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 7) return launch_rbgs_stream<2, 4, 128>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 10) return launch_rbgs_lean<8, 256>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 11) return launch_rbgs_lean<16, 256>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 12) return launch_rbgs_lean<32, 256>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 13) return launch_rbgs_lean<16, 512>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 14) return launch_rbgs_lean<8, 128>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 30) return launch_rbgs_col<8, 6, true, 2>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 31) return launch_rbgs_col<8, 5, true, 3>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 32) return launch_rbgs_col<8, 8, true, 2>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 33) return launch_rbgs_col<16, 6, true, 1>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 34) return launch_rbgs_col<8, 6, false, 2>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 35) return launch_rbgs_col<8, 4, true, 3>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 36) return launch_rbgs_col<16, 8, true, 1>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 37) return launch_rbgs_col<8, 5, false, 3>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 20) return launch_rbgs_stream<2, 8, 256>(sm_count, g, c, up, fp, op, omega, s);   // generic multi-stage kernel: 65-68 %
            // default: register-carried pair-column kernel (evo_kernels_rbcol.cuh), 97 % of the measured HBM peak at 513^3
            // (16-row tiles, 8-slot ring, one CTA per SM); smaller grids need more CTAs than 16-row tiles give
            if (g.n >= 385) return launch_rbgs_col<16, 8, true, 1>(sm_count, g, c, up, fp, op, omega, s);
            return launch_rbgs_col<8, 5, true, 3>(sm_count, g, c, up, fp, op, omega, s);
        }
        if (sweeps == 2) {
            if (variant == 1) return launch_rbgs_stream<4, 32, 512>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 2) return launch_rbgs_stream<4, 32, 256>(sm_count, g, c, up, fp, op, omega, s);
            if (variant == 3) return launch_rbgs_stream<4, 24, 512>(sm_count, g, c, up, fp, op, omega, s);
            return launch_rbgs_stream<4, 16, 256>(sm_count, g, c, up, fp, op, omega, s);
        }
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
// 3-D residual r = f - A u, 7-point star, one warp per row (lane-strided over x).  A CTA of 8 warps
// takes 8 consecutive rows of one plane so that the y-neighbour rows hit L1.  With NORM the squared
// residual is reduced per row in the canonical order (lane (x-1) mod 32 accumulates ascending x, then
// the xor butterfly) and written to rows[]; with !STORE the residual itself is never written
// (16 B/DOF: the residual that only feeds the solver's convergence test).
template <bool STORE, bool NORM>
__global__ void __launch_bounds__(256) k3_residual_rows(const Geom g, const Star7 c, const double *__restrict__ u,
                                                        const double *__restrict__ f, double *__restrict__ r,
                                                        double *__restrict__ rows)
{
    const int ni = g.n - 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y = 1 + blockIdx.x * 8 + warp, z = g.zlo + blockIdx.y;
    if (y > ni) return;
    const long long base = (long long)z * g.plane + (long long)y * g.pitch;
    const double *uc = u + base, *fc = f + base;
    const double *uym = uc - g.pitch, *uyp = uc + g.pitch, *uzm = uc - g.plane, *uzp = uc + g.plane;
    double acc = 0.0;
    // four x-strides of loads in flight per lane before the first use (the kernel is latency bound otherwise);
    // the sums and the canonical accumulation stay in ascending x order
    int x = 1 + lane;
    for (; x + 96 <= ni; x += 128) {
        double vzm[4], vym[4], vxm[4], vc[4], vxp[4], vyp[4], vzp[4], vf[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xx = x + 32 * k;
            vzm[k] = uzm[xx]; vym[k] = uym[xx]; vxm[k] = uc[xx - 1]; vc[k] = uc[xx]; vxp[k] = uc[xx + 1];
            vyp[k] = uyp[xx]; vzp[k] = uzp[xx]; vf[k] = fc[xx];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double sum = 0.0;
            sum = sum + c.zm * vzm[k];
            sum = sum + c.ym * vym[k];
            sum = sum + c.xm * vxm[k];
            sum = sum + c.c * vc[k];
            sum = sum + c.xp * vxp[k];
            sum = sum + c.yp * vyp[k];
            sum = sum + c.zp * vzp[k];
            const double rv = vf[k] - sum;
            if (STORE) r[base + x + 32 * k] = rv;
            if (NORM) acc = acc + rv * rv;
        }
    }
    for (; x <= ni; x += 32) {
        double sum = 0.0;
        sum = sum + c.zm * uzm[x];
        sum = sum + c.ym * uym[x];
        sum = sum + c.xm * uc[x - 1];
        sum = sum + c.c * uc[x];
        sum = sum + c.xp * uc[x + 1];
        sum = sum + c.yp * uyp[x];
        sum = sum + c.zp * uzp[x];
        const double rv = fc[x] - sum;
        if (STORE) r[base + x] = rv;
        if (NORM) acc = acc + rv * rv;
    }
    if (NORM) {
        acc = warp_butterfly(acc);
        if (lane == 0) rows[(long long)(z - g.zlo) * ni + (y - 1)] = acc;
    }
}

// 3-D pointwise weighted Jacobi sweep (`solve locally ... with jacobi`), same row mapping: reads the
// current slot, writes the next slot; per node s = sum of the off-diagonal terms in table order,
// x = (f - s) * (1/a), u_new = u + w (x - u)
static __global__ void __launch_bounds__(256) k3_jacobi_rows(const Geom g, const Star7 c, const double inv_c, const double omega,
                                                      const double *__restrict__ u, const double *__restrict__ f,
                                                      double *__restrict__ unew)
{
    const int ni = g.n - 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y = 1 + blockIdx.x * 8 + warp, z = g.zlo + blockIdx.y;
    if (y > ni) return;
    const long long base = (long long)z * g.plane + (long long)y * g.pitch;
    const double *uc = u + base, *fc = f + base;
    const double *uym = uc - g.pitch, *uyp = uc + g.pitch, *uzm = uc - g.plane, *uzp = uc + g.plane;
    int x = 1 + lane;
    for (; x + 96 <= ni; x += 128) {   // four x-strides of loads in flight per lane (see k3_residual_rows)
        double vzm[4], vym[4], vxm[4], vc[4], vxp[4], vyp[4], vzp[4], vf[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xx = x + 32 * k;
            vzm[k] = uzm[xx]; vym[k] = uym[xx]; vxm[k] = uc[xx - 1]; vc[k] = uc[xx]; vxp[k] = uc[xx + 1];
            vyp[k] = uyp[xx]; vzp[k] = uzp[xx]; vf[k] = fc[xx];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double sum = 0.0;
            sum = sum + c.zm * vzm[k];
            sum = sum + c.ym * vym[k];
            sum = sum + c.xm * vxm[k];
            sum = sum + c.xp * vxp[k];
            sum = sum + c.yp * vyp[k];
            sum = sum + c.zp * vzp[k];
            const double xs = (vf[k] - sum) * inv_c;
            unew[base + x + 32 * k] = vc[k] + omega * (xs - vc[k]);
        }
    }
    for (; x <= ni; x += 32) {
        double sum = 0.0;
        sum = sum + c.zm * uzm[x];
        sum = sum + c.ym * uym[x];
        sum = sum + c.xm * uc[x - 1];
        sum = sum + c.xp * uc[x + 1];
        sum = sum + c.yp * uyp[x];
        sum = sum + c.zp * uzp[x];
        const double xs = (fc[x] - sum) * inv_c;
        const double old = uc[x];
        unew[base + x] = old + omega * (xs - old);
    }
}

// 3-D trilinear prolongation + correction u += w * P e, one thread per coarse cell: the 8 coarse corner
// values give the 8 fine nodes (2X..2X+1, 2Y..2Y+1, 2Z..2Z+1); fine rows are updated with coalesced
// 16-byte read-modify-writes.  Per fine node the terms are added in ascending stencil-table order of
// the offsets o with x+o even (the order of the generic kernel and of the oracle).
struct DenseW { double w[27]; };

template <int PX, int PY, int PZ>
__device__ __forceinline__ double prolong_node(const DenseW &P, const double (&e)[2][2][2])
{
    // node parity (PX,PY,PZ): an even coordinate takes o = 0 from corner index 0, an odd one o = -1 (corner 0)
    // then o = +1 (corner 1)
    double acc = 0.0;
#pragma unroll
    for (int oz = -1; oz <= 1; ++oz) {
        if ((PZ == 0) != (oz == 0)) continue;
#pragma unroll
        for (int oy = -1; oy <= 1; ++oy) {
            if ((PY == 0) != (oy == 0)) continue;
#pragma unroll
            for (int ox = -1; ox <= 1; ++ox) {
                if ((PX == 0) != (ox == 0)) continue;
                acc = acc + P.w[(oz + 1) * 9 + (oy + 1) * 3 + (ox + 1)] * e[oz > 0][oy > 0][ox > 0];
            }
        }
    }
    return acc;
}

static __global__ void __launch_bounds__(128) k3_prolong_add(const Geom gf, const Geom gc, const DenseW P,
                                                      const double *__restrict__ ec, double *__restrict__ u,
                                                      const double weight, const int zc0)
{
    // coarse cell index in local coarse planes; the fine plane of local coarse plane Z is 2*(Z + gc.zoff) - gf.zoff
    const int X = blockIdx.x * 128 + threadIdx.x, Y = blockIdx.y, Z = zc0 + blockIdx.z;
    if (X > gc.n - 2) return;
    const int zf0 = 2 * (Z + gc.zoff) - gf.zoff;
    double e[2][2][2];
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const double *row = ec + (long long)(Z + dz) * gc.plane + (long long)(Y + dy) * gc.pitch + X;
            e[dz][dy][0] = row[0];
            e[dz][dy][1] = row[1];
        }
    const int n2 = gf.n - 2;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz) {
        const int z = zf0 + dz;
        if (z < gf.zlo || z > gf.zhi) continue;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int y = 2 * Y + dy;
            if (y < 1 || y > n2) continue;
            double2 *up = reinterpret_cast<double2 *>(u + (long long)z * gf.plane + (long long)y * gf.pitch + 2 * X);
            double2 v = *up;
            double p0, p1;
            if (dz == 0 && dy == 0) { p0 = prolong_node<0, 0, 0>(P, e); p1 = prolong_node<1, 0, 0>(P, e); }
            else if (dz == 0 && dy == 1) { p0 = prolong_node<0, 1, 0>(P, e); p1 = prolong_node<1, 1, 0>(P, e); }
            else if (dz == 1 && dy == 0) { p0 = prolong_node<0, 0, 1>(P, e); p1 = prolong_node<1, 0, 1>(P, e); }
            else { p0 = prolong_node<0, 1, 1>(P, e); p1 = prolong_node<1, 1, 1>(P, e); }
            if (X > 0) v.x = v.x + weight * p0;   // fine x = 0 is the boundary layer
            v.y = v.y + weight * p1;              // fine x = 2X+1 <= n-2 always
            *up = v;
        }
    }
}

template <typename T, int DIM, int NF>
static bool try_residual(int, const Geom &g, const OpSten &st, Fields<T> u, Fields<T> f, Fields<T> r, cudaStream_t s)
{
    if constexpr (std::is_same<T, double>::value && DIM == 3 && NF == 1) {
        Star7 c;
        if (g.n < 33 || !match_star7(st.s[0][0], &c)) return false;
        const int ni = g.n - 2;
        k3_residual_rows<true, false><<<dim3((ni + 7) / 8, g.zhi - g.zlo + 1), 256, 0, s>>>(g, c, u.p[0], f.p[0], r.p[0], nullptr);
        return cudaGetLastError() == cudaSuccess;
    } else {
        return false;
    }
}

// residual + canonical row sums of |r|^2 (rows[(z-1)*ni + (y-1)]); store = also write the residual field
template <typename T, int DIM, int NF>
static bool try_residual_norm(const Geom &g, const OpSten &st, Fields<T> u, Fields<T> f, Fields<T> r, double *rows,
                              bool store, cudaStream_t s)
{
    if constexpr (std::is_same<T, double>::value && DIM == 3 && NF == 1) {
        Star7 c;
        if (g.n < 33 || !match_star7(st.s[0][0], &c)) return false;
        const int ni = g.n - 2;
        if (store) k3_residual_rows<true, true><<<dim3((ni + 7) / 8, g.zhi - g.zlo + 1), 256, 0, s>>>(g, c, u.p[0], f.p[0], r.p[0], rows);
        else k3_residual_rows<false, true><<<dim3((ni + 7) / 8, g.zhi - g.zlo + 1), 256, 0, s>>>(g, c, u.p[0], f.p[0], r.p[0], rows);
        return cudaGetLastError() == cudaSuccess;
    } else {
        return false;
    }
}

template <typename T, int DIM, int NF>
static bool try_smooth_point(int, const Geom &g, const OpSten &st, const SmoothParams &sp, Fields<T> src, Fields<T> dst,
                             Fields<T> rhs, cudaStream_t s)
{
    if constexpr (std::is_same<T, double>::value && DIM == 3 && NF == 1) {
        Star7 c;
        // only the out-of-place Jacobi sweep (colour passes of RB-GS go through the streaming kernel)
        if (sp.color >= 0 || src.p[0] == dst.p[0] || g.n < 33 || !match_star7(st.s[0][0], &c)) return false;
        const int ni = g.n - 2;
        if (g.zhi < g.zlo) return true;
        k3_jacobi_rows<<<dim3((ni + 7) / 8, g.zhi - g.zlo + 1), 256, 0, s>>>(g, c, 1.0 / c.c, sp.omega, src.p[0], rhs.p[0], dst.p[0]);
        return cudaGetLastError() == cudaSuccess;
    } else {
        return false;
    }
}
// ---------------------------------------------------------------------------------------------
// Fused RHS@(l-1) = R (f - A u), 3-D 7-point operator, dense 27-point restriction: the fine residual never
// reaches HBM (16 B/fine DOF read + 1 B written instead of 24 + 9 for the two separate statements).
// A CTA owns CY x CX coarse nodes and streams over a chunk of coarse planes.  u and f planes of the fine tile
// (+ halo) arrive by TMA into shared-memory rings, issued a full step (two fine planes) ahead, so HBM stays
// busy while the CTA computes.  Per coarse plane Z the two new fine residual planes 2Z, 2Z+1 are computed from
// shared memory into a 5-slot residual ring (plane 2Z-1 is kept from the previous step; x-even and x-odd columns
// are stored separately so that the stride-2 reads of the restriction are conflict free), then the coarse
// values are formed by adding the 27 terms in ascending stencil-table order like the generic kernel.  Per node the
// arithmetic is that of k3_residual_rows / k_restrict -> bit-identical.  The residual is 0 on the boundary layer
// (only x = 0 / n-1 can be touched: 2Y+-1 and 2Z+-1 are inner for inner coarse nodes).  z-slab aware.
template <int CY, int CX>
struct RrCfg {
    static constexpr int RROWS = 2 * CY + 1;                       // fine residual rows of the tile
    static constexpr int NPAIR = CX + 1;                           // column pairs (fine x = 2X0-2+2j, +1), j = 0 .. CX
    static constexpr int ITEMS = RROWS * NPAIR;                    // one thread per (row, pair), fixed for the whole stream
    static constexpr int RT = (ITEMS + 31) / 32 * 32;              // residual threads (producer warps)
    static constexpr int CT = CY * CX;                             // restriction threads (consumer warps), one coarse node each
    static constexpr int NT = RT + CT;
    static constexpr int LX = 2 * CX + 4;                          // box width: fine x = 2*X0-2 .. 2*X0+2*CX+1 (even start)
    static constexpr int ULY = 2 * CY + 3, FLY = RROWS;            // box rows of u (halo) and f
    static constexpr int NU = 6, NF = 4, NR = 5;                   // ring depths (planes)
    static constexpr int USTRIDE = (LX * ULY * 8 + 127) / 128 * 16, FSTRIDE = (LX * FLY * 8 + 127) / 128 * 16;   // doubles
    static constexpr int RA = CX + 1, RB = CX;                     // odd-x (2X-1, first) / even-x (2X, at offset RA) entries per row
    static constexpr int RPITCH = RA + RB + 1;                     // doubles per residual row
    static constexpr int RSTRIDE = RROWS * RPITCH;
    static constexpr size_t SMEM = ((size_t)NU * USTRIDE + (size_t)NF * FSTRIDE + (size_t)NR * RSTRIDE) * 8;
    static_assert(CT % 32 == 0, "whole consumer warps");
};

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Warp-specialised: RT producer threads compute residual planes (one (row, column pair) each, the centre values of
// planes p-1, p, p+1 stay in registers while streaming), CT consumer threads form the coarse values.  The
// 27-term dependent sum of a coarse node (~500 cycles of latency) then overlaps the producers' next step.
// Named barriers: 1 = producers only (u/f ring slots dead -> next TMA), 2/3 = residual planes of step k ready
// (producers arrive, consumers wait), 4/5 = consumers finished step k (producers wait before step k+2 reuses the
// residual ring slots).
template <int CY, int CX>
__global__ void __launch_bounds__(RrCfg<CY, CX>::NT) k3_residual_restrict_tma(const __grid_constant__ CUtensorMap umap,
                                                                             const __grid_constant__ CUtensorMap fmap,
                                                                             const Geom gf, const Geom gc, const Star7 c,
                                                                             const DenseW R, double *__restrict__ dst,
                                                                             const int zchunk)
{
    using C = RrCfg<CY, CX>;
    extern __shared__ __align__(128) double rr_smem[];
    __shared__ __align__(8) uint64_t ubar[C::NU], fbar[C::NF];
    double *uring = rr_smem, *fring = rr_smem + (size_t)C::NU * C::USTRIDE, *rring = fring + (size_t)C::NF * C::FSTRIDE;
    const int tid = threadIdx.x;
    const int X0 = 1 + blockIdx.x * CX, Y0 = 1 + blockIdx.y * CY;
    const int Za = gc.zlo + blockIdx.z * zchunk, Zb = min(Za + zchunk - 1, gc.zhi);
    const int nsteps = Zb - Za + 1;
    const int nfi = gf.n - 2, nci = gc.n - 2;
    const int zr0 = 2 * (Za + gc.zoff) - gf.zoff - 1;                // first fine residual plane (local index)

    if (tid == 0) {
        for (int i = 0; i < C::NU; ++i) mbar_init(&ubar[i], 1);
        for (int i = 0; i < C::NF; ++i) mbar_init(&fbar[i], 1);
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= C::RT) {
        // ---------------- consumers: coarse node (X, Y) of every coarse plane of the chunk ----------------
        const int t = tid - C::RT;
        const int ry = t / CX, rx = t - ry * CX;
        const int X = X0 + rx, Y = Y0 + ry;
        const bool ok = X <= nci && Y <= nci;
        const int o = (2 * ry) * C::RPITCH + rx;
        int sm = zr0 % C::NR;                                         // ring slot of fine plane 2Z-1
        double *out = dst + (long long)Za * gc.plane + (long long)Y * gc.pitch + X;
        for (int k = 0; k < nsteps; ++k) {
            const int s0 = sm + 1 >= C::NR ? sm + 1 - C::NR : sm + 1, sp = sm + 2 >= C::NR ? sm + 2 - C::NR : sm + 2;
            named_bar_sync(2 + (k & 1), C::NT);
            if (ok) {
                double acc = 0.0;
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {
                    const double *pl = rring + (dz == 0 ? sm : (dz == 1 ? s0 : sp)) * C::RSTRIDE + o;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const double *row = pl + dy * C::RPITCH;
                        acc = acc + R.w[dz * 9 + dy * 3 + 0] * row[0];           // fine x = 2X - 1
                        acc = acc + R.w[dz * 9 + dy * 3 + 1] * row[C::RA];       // fine x = 2X
                        acc = acc + R.w[dz * 9 + dy * 3 + 2] * row[1];           // fine x = 2X + 1
                    }
                }
                *out = acc;
            }
            out += gc.plane;
            sm = sp;                                                  // plane 2Z+1 is plane 2(Z+1)-1
            if (k + 2 < nsteps) named_bar_arrive(4 + (k & 1), C::NT);
        }
        return;
    }

    // ---------------- producers ----------------
    const int xb = 2 * X0 - 2, yub = 2 * Y0 - 2, yfb = 2 * Y0 - 1;   // box origins (fine coordinates)
    const int zr1 = zr0 + 2 * nsteps;                                 // last fine residual plane
    const int q0 = zr0 - 1, qmax = zr1 + 1;                           // u planes q0 .. qmax, f planes zr0 .. zr1
    int u_issued = q0 - 1, f_issued = zr0 - 1;                        // only thread 0 uses these
    auto issue_u = [&](int upto) {
        for (upto = min(upto, qmax); u_issued < upto;) {
            const int q = ++u_issued, sl = (q - q0) % C::NU;
            mbar_expect_tx(&ubar[sl], C::LX * C::ULY * 8);
            tma_load_plane(uring + (size_t)sl * C::USTRIDE, &umap, xb, yub, q, &ubar[sl]);
        }
    };
    auto issue_f = [&](int upto) {
        for (upto = min(upto, zr1); f_issued < upto;) {
            const int q = ++f_issued, sl = (q - zr0) % C::NF;
            mbar_expect_tx(&fbar[sl], C::LX * C::FLY * 8);
            tma_load_plane(fring + (size_t)sl * C::FSTRIDE, &fmap, xb, yfb, q, &fbar[sl]);
        }
    };
    if (tid == 0) {
        issue_u(q0 + C::NU - 1);
        issue_f(zr0 + C::NF - 1);
    }

    // this thread's residual item: row r, column pair j -> fine x = xa = 2(X0+j-1) (even-x array entry j-1, at
    // offset RA) and xa + 1 = 2(X0+j)-1 (odd-x array entry j); the pair is 16-byte aligned in the box (column 2j)
    const bool active = tid < C::ITEMS;
    const int r = active ? tid / C::NPAIR : 0, j = active ? tid - (tid / C::NPAIR) * C::NPAIR : 0;
    const int xa = xb + 2 * j, yy = yfb + r;
    const bool ok_a = active && j >= 1 && xa <= nfi && yy <= nfi;       // j = 0: x = 2X0-2 belongs to the left neighbour
    const bool ok_b = active && xa + 1 <= nfi && yy <= nfi;
    const int ou = (r + 1) * C::LX + 2 * j, of = r * C::LX + 2 * j;
    const int oxm = ou - (j >= 1 ? 1 : 0);                               // left neighbour of the pair (unused for j = 0)
    const int res_a = r * C::RPITCH + C::RA + (j >= 1 ? j - 1 : 0), res_b = r * C::RPITCH + j;
    double2 um = make_double2(0.0, 0.0), u0 = um;

    // ring cursors: slot / phase of the next u plane (p+1) and f plane (p) to wait for, slots of planes p and residual p
    int su = 1, sun = 2, sun_ph = 0, sf = 0, sf_ph = 0, sr = zr0 % C::NR;
    // residual of the next fine plane p for this thread's pair; um / u0 hold the centre pairs of planes p-1 / p
    auto residual_plane = [&]() {
        mbar_wait(&ubar[sun], (uint32_t)sun_ph);
        mbar_wait(&fbar[sf], (uint32_t)sf_ph);
        if (active) {
            const double *s0 = uring + su * C::USTRIDE + ou;
            const double2 up = *reinterpret_cast<const double2 *>(uring + sun * C::USTRIDE + ou);
            const double2 ym = *reinterpret_cast<const double2 *>(s0 - C::LX);
            const double2 yp = *reinterpret_cast<const double2 *>(s0 + C::LX);
            const double2 fv = *reinterpret_cast<const double2 *>(fring + sf * C::FSTRIDE + of);
            const double xm = s0[oxm - ou], xp = s0[2];
            double sa = 0.0, sb = 0.0;
            sa = sa + c.zm * um.x;  sb = sb + c.zm * um.y;
            sa = sa + c.ym * ym.x;  sb = sb + c.ym * ym.y;
            sa = sa + c.xm * xm;    sb = sb + c.xm * u0.x;
            sa = sa + c.c * u0.x;   sb = sb + c.c * u0.y;
            sa = sa + c.xp * u0.y;  sb = sb + c.xp * xp;
            sa = sa + c.yp * yp.x;  sb = sb + c.yp * yp.y;
            sa = sa + c.zp * up.x;  sb = sb + c.zp * up.y;
            double *rs = rring + sr * C::RSTRIDE;
            if (j >= 1) rs[res_a] = ok_a ? fv.x - sa : 0.0;
            rs[res_b] = ok_b ? fv.y - sb : 0.0;
            um = u0;
            u0 = up;
        }
        su = sun;
        if (++sun == C::NU) { sun = 0; sun_ph ^= 1; }
        if (++sf == C::NF) { sf = 0; sf_ph ^= 1; }
        if (++sr == C::NR) sr = 0;
    };

    // prologue: fine residual plane zr0 (= 2 Za - 1); planes q0, q0+1 feed the register window
    mbar_wait(&ubar[0], 0u);
    mbar_wait(&ubar[1], 0u);
    if (active) {
        um = *reinterpret_cast<const double2 *>(uring + ou);
        u0 = *reinterpret_cast<const double2 *>(uring + (size_t)C::USTRIDE + ou);
    }
    residual_plane();
    named_bar_sync(1, C::RT);
    if (tid == 0) { fence_proxy_async(); issue_u(zr0 + C::NU - 1); issue_f(zr0 + C::NF); }

    for (int k = 0; k < nsteps; ++k) {
        const int zf = zr0 + 1 + 2 * k;
        if (k >= 2) named_bar_sync(4 + (k & 1), C::NT);   // consumers are done with the residual ring slots of step k-2
        residual_plane();
        residual_plane();
        named_bar_arrive(2 + (k & 1), C::NT);             // residual planes zf, zf+1 are in the ring
        named_bar_sync(1, C::RT);                         // u planes <= zf and f planes <= zf+1 are dead
        if (tid == 0) { fence_proxy_async(); issue_u(zf + C::NU); issue_f(zf + 1 + C::NF); }
    }
}

template <int CY, int CX>
static bool launch_residual_restrict(int sm_count, const Geom &gf, const Geom &gc, const Star7 &c, const DenseW &W, const double *u,
                                     const double *f, double *dst, cudaStream_t s)
{
    using C = RrCfg<CY, CX>;
    CUtensorMap um, fm;
    if (!make_plane_map(&um, gf, u, C::LX, C::ULY) || !make_plane_map(&fm, gf, f, C::LX, C::FLY)) return false;
    static int occ = 0;
    if (occ == 0) {
        if (cudaFuncSetAttribute(k3_residual_restrict_tma<CY, CX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess)
            return false;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3_residual_restrict_tma<CY, CX>, C::NT, C::SMEM) != cudaSuccess || occ < 1) occ = 1;
    }
    const int nci = gc.n - 2, planes = gc.zhi - gc.zlo + 1;
    if (planes <= 0) return true;
    const int tx = (nci + CX - 1) / CX, ty = (nci + CY - 1) / CY;
    // z chunks: minimise (waves) x (coarse planes per chunk + pipeline fill); a chunk recomputes one fine plane
    const long long slots = (long long)occ * sm_count;
    int best = 1;
    double best_cost = 1e300;
    for (int ch = 1; ch <= 32 && (ch == 1 || ch * 4 <= planes); ++ch) {
        const int zc = (planes + ch - 1) / ch;
        const long long ctas = (long long)tx * ty * ((planes + zc - 1) / zc);
        const double cost = (double)((ctas + slots - 1) / slots) * (zc + 3);
        if (cost < best_cost) { best_cost = cost; best = ch; }
    }
    const int zchunk = (planes + best - 1) / best;
    const int chunks = (planes + zchunk - 1) / zchunk;
    k3_residual_restrict_tma<CY, CX><<<dim3(tx, ty, chunks), C::NT, C::SMEM, s>>>(um, fm, gf, gc, c, W, dst, zchunk);
    return cudaGetLastError() == cudaSuccess;
}

template <typename T, int DIM, int NF>
static bool try_residual_restrict(int sm_count, const Geom &gf, const Geom &gc, const OpSten &st, const TransferW &R, Fields<T> u,
                                  Fields<T> f, Fields<T> dst, cudaStream_t s)
{
    if constexpr (std::is_same<T, double>::value && DIM == 3 && NF == 1) {
        Star7 c;
        if (gf.n < 33 || R.nnz != 27 || !match_star7(st.s[0][0], &c) || get_encode_tiled() == nullptr) return false;
        DenseW W;
        for (int q = 0; q < 27; ++q) W.w[(R.oz[q] + 1) * 9 + (R.oy[q] + 1) * 3 + (R.ox[q] + 1)] = R.w[q];
        const int variant = option(OPT_RR_VARIANT);
        switch (variant) {
        case 1: return launch_residual_restrict<2, 64>(sm_count, gf, gc, c, W, u.p[0], f.p[0], dst.p[0], s);
        case 2: return launch_residual_restrict<4, 64>(sm_count, gf, gc, c, W, u.p[0], f.p[0], dst.p[0], s);
        case 3: return launch_residual_restrict<8, 32>(sm_count, gf, gc, c, W, u.p[0], f.p[0], dst.p[0], s);
        case 4: return launch_residual_restrict<2, 32>(sm_count, gf, gc, c, W, u.p[0], f.p[0], dst.p[0], s);
        default: return launch_residual_restrict<4, 32>(sm_count, gf, gc, c, W, u.p[0], f.p[0], dst.p[0], s);
        }
    } else {
        return false;
    }
}
template <typename T, int DIM, int NF>
static bool try_prolong_add(int, const Geom &gf, const Geom &gc, const TransferW &P, Fields<T> src, Fields<T> dst, double weight,
                            cudaStream_t s)
{
    if constexpr (std::is_same<T, double>::value && DIM == 3 && NF == 1) {
        if (gf.n < 33) return false;
        DenseW W;
        for (int i = 0; i < 27; ++i) W.w[i] = 0.0;
        for (int q = 0; q < P.nnz; ++q) W.w[(P.oz[q] + 1) * 9 + (P.oy[q] + 1) * 3 + (P.ox[q] + 1)] = P.w[q];
        const int cells = gc.n - 1;
        // coarse cells (local planes) whose two fine planes intersect the owned fine range
        const int zc0 = std::max(0, (gf.zlo + gf.zoff) / 2 - gc.zoff);
        const int zc1 = std::min(gc.nz - 2, (gf.zhi + gf.zoff) / 2 - gc.zoff);
        if (zc1 < zc0) return true;
        k3_prolong_add<<<dim3((cells + 127) / 128, cells, zc1 - zc0 + 1), 128, 0, s>>>(gf, gc, W, src.p[0], dst.p[0], weight, zc0);
        return cudaGetLastError() == cudaSuccess;
    } else {
        return false;
    }
}

}  // namespace star
}  // namespace evo
