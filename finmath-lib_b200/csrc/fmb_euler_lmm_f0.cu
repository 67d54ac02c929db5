// eulerLmmKernel instantiations for a run-time factor count (F > 8): see fmb_euler_lmm.cuh
#include "fmb_euler_lmm.cuh"
namespace fmb { FMB_LMM_DEFINE(0) }
