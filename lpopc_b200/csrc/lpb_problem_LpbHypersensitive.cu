// Kernel instantiation for the functor set LpbHypersensitive (include/problems/hypersensitive.h).
#include "../../include/problems/hypersensitive.h"
#include "lpb_hessian.cuh"
LPB_DEFINE_FUNCTOR(LpbHypersensitive)
