// Kernel instantiation for the functor set LpbSynthetic20 (include/problems/synthetic20.h).
#include "../../include/problems/synthetic20.h"
#include "lpb_hessian.cuh"
LPB_DEFINE_FUNCTOR(LpbSynthetic20)
