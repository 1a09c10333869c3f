// Kernel instantiation for the functor set LpbQuadrotor (include/problems/quadrotor.h).
#include "../../include/problems/quadrotor.h"
#include "lpb_hessian.cuh"
LPB_DEFINE_FUNCTOR(LpbQuadrotor)
