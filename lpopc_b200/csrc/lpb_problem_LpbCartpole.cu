// Kernel instantiation for the functor set LpbCartpole (include/problems/cartpole.h).
#include "../../include/problems/cartpole.h"
#include "lpb_hessian.cuh"
LPB_DEFINE_FUNCTOR(LpbCartpole)
