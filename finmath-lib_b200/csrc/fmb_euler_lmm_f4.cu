// eulerLmmKernel instantiations for F = 4 factors: see fmb_euler_lmm.cuh
#include "fmb_euler_lmm.cuh"
namespace fmb { FMB_LMM_DEFINE(4) }
