// Kernel instantiation for the functor set LpbLaunch (include/problems/launch.h).
#include "../../include/problems/launch.h"
#include "lpb_hessian.cuh"
LPB_DEFINE_FUNCTOR(LpbLaunch)
