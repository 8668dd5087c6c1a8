// evo_kernels_star.cuh -- specialised high-bandwidth kernels for scalar real star stencils
// (5-point 2-D / 7-point 3-D).  Every try_* function returns false when the statement does not match
// its fast path; the caller then launches the generic kernel.  Results are bit-identical to the
// generic kernels (same operation order, -fmad=false).
#pragma once
#include "evo_kernels.cuh"

namespace evo {
namespace star {

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
