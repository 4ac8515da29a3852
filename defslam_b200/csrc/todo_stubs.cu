/*
 * todo_stubs.cu -- entry points declared in include/defslam_b200.h whose CUDA
 * implementation has not landed yet.  They fail loudly (DEFSLAM_ENOTIMPL);
 * there is no CPU fallback behind them.  Each stub disappears when its kernel
 * is written.
 */
#include "../../include/defslam_b200.h"

extern "C" {
int defslam_schwarp_fit(const defslam_schwarp_problem *, defslam_diffprop *) { return DEFSLAM_ENOTIMPL; }
int defslam_schwarp_evaluate(const defslam_schwarp_problem *, double *, double *) { return DEFSLAM_ENOTIMPL; }
int defslam_normals_batched(const defslam_normals_problem *, double *, double *, float *, int32_t *) {
  return DEFSLAM_ENOTIMPL;
}
int defslam_sfn_solve(const defslam_sfn_problem *, double *, float *) { return DEFSLAM_ENOTIMPL; }
}
