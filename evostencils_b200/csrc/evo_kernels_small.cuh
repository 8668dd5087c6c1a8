// evo_kernels_small.cuh -- pointwise red-black Gauss-Seidel on a tiny grid (<= 4096 inner nodes), 5-/7-point star.
//
// Statement: `color with { (i0+i1+i2)%2 } solve locally at u@l relax w { u@l => A@l * u@l == f@l }`
// (evostencils/code_generation/exastencils.py:659-682, :769-822) on the latency-bound levels (3-D <= 17^3, 2-D <= 65^2):
// they are visited 2^k times per W-cycle and by every individual of a generation.  k_smooth_rb_small (generic
// local_solve on global memory) spends ~3.5 us per colour pass; here u and f are staged ONCE into shared memory
// (compact layout), ALL merged repetitions and both colours run on them with straight-line star arithmetic, and the inner
// nodes are written back -- one launch of a few microseconds.  Per node exactly the arithmetic of the oracle's
// sweep_pointwise_scalar (and of k3_rbgs_col / k2_sweep_warp):
//   s = sum of the off-diagonal terms in ascending table order, x = (f - s) * (1/a), u += w (x - u)    -> bit-identical.
#pragma once
#include "evo_kernels_star.cuh"
#include "evo_kernels_warp2d.cuh"

namespace evo {
namespace small {

template <int DIM>
__global__ void __launch_bounds__(1024) k_star_rb_small(const Geom g, const star::Star7 c /* 2-D: zm, zp unused */, const double inv_c,
                                                        const double omega, double *__restrict__ u, const double *__restrict__ f,
                                                        const int sweeps)
{
    extern __shared__ double small_sm[];
    const int n = g.n, ni = n - 2, nzi = DIM == 3 ? ni : 1;
    const int vol = n * n * (DIM == 3 ? n : 1);
    double *su = small_sm, *sf = small_sm + vol;
    for (int t = threadIdx.x; t < vol; t += 1024) {
        const int x = t % n, r = t / n, y = r % n, z = r / n;
        const long long gi = node_index(g, x, y, z);
        su[t] = u[gi];
        sf[t] = f[gi];
    }
    __syncthreads();
    const int half = (ni + 1) / 2, total = half * ni * nzi;
    for (int sw = 0; sw < sweeps; ++sw)
        for (int color = 0; color < 2; ++color) {
            for (int t = threadIdx.x; t < total; t += 1024) {
                const int tx = t % half, row = t / half;
                const int y = 1 + row % ni, z = DIM == 3 ? 1 + row / ni : 0;
                const int x = 1 + 2 * tx + ((1 + y + z + color) & 1);
                if (x > ni) continue;
                const int i = (z * n + y) * n + x;
                double sum = 0.0;
                if (DIM == 3) sum = sum + c.zm * su[i - n * n];
                sum = sum + c.ym * su[i - n];
                sum = sum + c.xm * su[i - 1];
                sum = sum + c.xp * su[i + 1];
                sum = sum + c.yp * su[i + n];
                if (DIM == 3) sum = sum + c.zp * su[i + n * n];
                const double xs = (sf[i] - sum) * inv_c;
                const double old = su[i];
                su[i] = old + omega * (xs - old);
            }
            __syncthreads();
        }
    const int inner = ni * ni * nzi;
    for (int t = threadIdx.x; t < inner; t += 1024) {
        const int x = 1 + t % ni, r = t / ni, y = 1 + r % ni, z = DIM == 3 ? 1 + r / ni : 0;
        u[node_index(g, x, y, z)] = su[(z * n + y) * n + x];
    }
}

// `sweeps` in-place red-black sweeps of field 0; false = not applicable (the caller falls back to the generic kernel)
template <typename T, int DIM, int NF>
static bool try_rb_small(const Geom &g, const OpSten &st, Fields<T> u, Fields<T> f, double omega, int sweeps, cudaStream_t s)
{
    if constexpr (std::is_same<T, double>::value && NF == 1) {
        const long long ni = g.n - 2;
        if (ni * ni * (DIM == 3 ? ni : 1) > 4096) return false;
        star::Star7 c;
        if (DIM == 3) {
            if (!star::match_star7(st.s[0][0], &c)) return false;
        } else {
            w2::Star5 c5;
            if (!w2::match_star5(st.s[0][0], &c5)) return false;
            c.zm = c.zp = 0.0; c.ym = c5.ym; c.xm = c5.xm; c.c = c5.c; c.xp = c5.xp; c.yp = c5.yp;
        }
        const size_t smem = 2 * (size_t)g.n * g.n * (DIM == 3 ? g.n : 1) * sizeof(double);
        if (smem > 160 * 1024) return false;
        static bool attr = false;
        if (!attr) {
            if (cudaFuncSetAttribute(k_star_rb_small<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess) return false;
            attr = true;
        }
        k_star_rb_small<DIM><<<1, 1024, smem, s>>>(g, c, 1.0 / c.c, omega, u.p[0], f.p[0], sweeps);
        return cudaGetLastError() == cudaSuccess;
    } else {
        return false;
    }
}

}  // namespace small
}  // namespace evo
