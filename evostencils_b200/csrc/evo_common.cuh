// evo_common.cuh -- shared device/host helpers of the B200 multigrid evaluation library.
// Compiled only for sm_100a (nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false).
// -fmad=false: the CPU oracle is built with -ffp-contract=off; with identical operation order the
// pointwise kernels are bit-identical to it, which is what the parity tests assert.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <cstring>

#include "../../include/evostencils_b200.h"

namespace evo {

// ---------------------------------------------------------------------------------------------
// Tuning switches (kernel variants, experiments).  Each one is initialised from its environment variable on
// first use and can be changed at run time through evo_set_option() -- a new value takes effect for cycles
// whose solver graph is captured afterwards.  Defaults are the measured best.
enum {
    OPT_RB_VARIANT,      // EVO_RB_VARIANT      tile shapes / kernel of the 3-D RB-GS sweep
    OPT_RB_FUSE2,        // EVO_RB_FUSE2        two consecutive RB-GS sweeps per launch (temporal blocking)
    OPT_RR_VARIANT,      // EVO_RR_VARIANT      tiles of the fused residual + restriction
    OPT_NO_PINGPONG,     // EVO_NO_PINGPONG     copy the [next] slot back every cycle
    OPT_CG_GLOBAL,       // EVO_CG_GLOBAL       coarse CG with its vectors in global memory
    OPT_CG_NOREG,        // EVO_CG_NOREG        no register-resident coarse CG
    OPT_ROWSEQ_GLOBAL,   // EVO_ROWSEQ_GLOBAL   order-dependent coloured sweeps without the shared-memory window
    OPT_ROWSEQ_NOPIPE,   // EVO_ROWSEQ_NOPIPE   ... without the pipelined passes
    OPT_FAS_CGS,         // EVO_FAS_CGS         FAS coarse solver: 0 auto, 1 one CTA, 2 one launch per sweep
    OPT_FAS_CGS_GLOBAL,  // EVO_FAS_CGS_GLOBAL  FAS coarse solver without shared memory
    OPT_LEX_VARIANT,     // EVO_LEX_VARIANT     lexicographic sweeps: 0 cluster wavefront, 1 single CTA
    OPT_STAR2D,          // EVO_STAR2D          0 = generic 2-D kernels only (no specialised 5-point path)
    OPT_COARSE_FUSE,     // EVO_COARSE_FUSE     fused runs on small levels: 0 off (default: measured slower), 1 one CTA, 2 also clusters, 3 smallest only
    OPT_NO_ZERO_FUSE,    // EVO_NO_ZERO_FUSE    keep `SOL@(l-1) = 0` as its own node instead of folding it into the restriction before it
    OPT_RR_WARP_NODES,   // EVO_RR_WARP_NODES   generic fused residual+restriction: warp-per-coarse-node kernel up to this many coarse nodes (0 = default)
    OPT_COUNT
};
struct OptionTable {
    int value[OPT_COUNT];
    bool init[OPT_COUNT];
};
inline const char *option_name(int id)
{
    static const char *names[OPT_COUNT] = {"EVO_RB_VARIANT", "EVO_RB_FUSE2", "EVO_RR_VARIANT", "EVO_NO_PINGPONG", "EVO_CG_GLOBAL",
                                           "EVO_CG_NOREG", "EVO_ROWSEQ_GLOBAL", "EVO_ROWSEQ_NOPIPE", "EVO_FAS_CGS",
                                           "EVO_FAS_CGS_GLOBAL", "EVO_LEX_VARIANT", "EVO_STAR2D", "EVO_COARSE_FUSE", "EVO_NO_ZERO_FUSE", "EVO_RR_WARP_NODES"};
    return names[id];
}
inline OptionTable &option_table()
{
    static OptionTable t = {};
    return t;
}
inline int option_default(int id) { return id == OPT_STAR2D ? 1 : 0; }
inline int option(int id)
{
    OptionTable &t = option_table();
    if (!t.init[id]) {
        const char *e = getenv(option_name(id));
        // a variable that is set but empty / not a number counts as 1 (the historical "is it set" switches)
        t.value[id] = e ? ((*e >= '0' && *e <= '9') || *e == '-' ? atoi(e) : 1) : option_default(id);
        t.init[id] = true;
    }
    return t.value[id];
}

// ---------------------------------------------------------------------------------------------
// complex fp64 as one 16-byte word (double2): one Helmholtz unknown per LDG.128/STG.128
struct __align__(16) cplx {
    double re, im;
    __host__ __device__ cplx() : re(0.0), im(0.0) {}
    __host__ __device__ cplx(double r) : re(r), im(0.0) {}
    __host__ __device__ cplx(double r, double i) : re(r), im(i) {}
};
__host__ __device__ inline cplx operator+(cplx a, cplx b) { return cplx(a.re + b.re, a.im + b.im); }
__host__ __device__ inline cplx operator-(cplx a, cplx b) { return cplx(a.re - b.re, a.im - b.im); }
__host__ __device__ inline cplx operator*(cplx a, cplx b) { return cplx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__host__ __device__ inline cplx operator*(double a, cplx b) { return cplx(a * b.re, a * b.im); }
__host__ __device__ inline cplx operator/(cplx a, cplx b)
{
    // Smith's algorithm (what libgcc's __divdc3 does for finite operands)
    if (fabs(b.re) < fabs(b.im)) {
        double ratio = b.re / b.im, denom = b.re * ratio + b.im;
        return cplx((a.re * ratio + a.im) / denom, (a.im * ratio - a.re) / denom);
    }
    double ratio = b.im / b.re, denom = b.im * ratio + b.re;
    return cplx((a.im * ratio + a.re) / denom, (a.im - a.re * ratio) / denom);
}
__host__ __device__ inline double abs2(double a) { return a * a; }
__host__ __device__ inline double abs2(cplx a) { return a.re * a.re + a.im * a.im; }

template <typename T> struct scalar_traits;
template <> struct scalar_traits<double> {
    static constexpr int words = 1;
    __host__ __device__ static double make(double re, double) { return re; }
};
template <> struct scalar_traits<cplx> {
    static constexpr int words = 2;
    __host__ __device__ static cplx make(double re, double im) { return cplx(re, im); }
};

// ---------------------------------------------------------------------------------------------
// geometry of one level: (2^l + 1)^dim nodes, x fastest, rows padded to a multiple of 16 entries
// (128 B for fp64) so that every row starts on a cache-line / TMA-legal boundary.
struct Geom {
    int n;          // nodes per dimension
    int dim;        // 2 or 3
    int pitch;      // entries per row (>= n, multiple of 16)
    int nz;         // n for 3-D, 1 for 2-D
    long long plane;  // pitch * n
    long long total;  // plane * nz
    // z-slab view (domain decomposition; for an undecomposed grid: zlo = zin0 = 1, zhi = zin1 = n-2, zpar = 0,
    // zoff = 0): the array holds planes [zoff, zoff + nz) of the global grid
    int zlo, zhi;     // local plane range this rank owns / sweeps over
    int zin0, zin1;   // local range of planes that are inner planes of the GLOBAL grid (halo recomputation limit)
    int zpar;         // parity of zoff: global colour of a node = (x + y + z_local + zpar) & 1
    int zoff;         // global z index of local plane 0
};

// stencil of one (row field, column field) block, non-zeros in ascending table order
struct Sten {
    int nnz;
    signed char ox[27], oy[27], oz[27];
    double re[27], im[27];
};
struct OpSten {  // passed by value as a __grid_constant__ kernel parameter (constant bank reads)
    Sten s[EVO_MAX_FIELDS][EVO_MAX_FIELDS];
};

template <typename T> struct Fields {  // per-field array pointers of one buffer kind
    T *p[EVO_MAX_FIELDS];
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block reduction: warp shuffles, then warp 0 adds the per-warp values in order
__device__ __forceinline__ double block_sum(double v, double *smem /* >= 32 doubles */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = lane < nwarp ? smem[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;  // valid in warp 0
}

}  // namespace evo
