// evo_dispatch_inst.cu -- explicit instantiation of the statement dispatch for ONE (scalar type, dimension, fields)
// combination, selected with -DEVO_INST=0..4; the build compiles the five combinations in parallel.
#include "evo_dispatch.cuh"

#ifndef EVO_INST
#error "compile with -DEVO_INST=0..4"
#endif
#if EVO_INST == 0
#define EVO_INST_ARGS double, 2, 1
#elif EVO_INST == 1
#define EVO_INST_ARGS double, 2, 2
#elif EVO_INST == 2
#define EVO_INST_ARGS double, 3, 1
#elif EVO_INST == 3
#define EVO_INST_ARGS double, 3, 2
#else
#define EVO_INST_ARGS cplx, 2, 1
#endif

template int enqueue_op<EVO_INST_ARGS>(evo_cycle *, const evo_op &, cudaStream_t);
template int op_residual<EVO_INST_ARGS>(evo_cycle *, int, bool, cudaStream_t);
template int op_restrict<EVO_INST_ARGS>(evo_cycle *, const evo_op &, cudaStream_t);
template int op_reduce_rows<EVO_INST_ARGS>(evo_cycle *, int, cudaStream_t);
template bool run_eligible<EVO_INST_ARGS>(const evo_cycle *, const evo_op &);
template int enqueue_run<EVO_INST_ARGS>(evo_cycle *, const evo_op *, int, cudaStream_t);
