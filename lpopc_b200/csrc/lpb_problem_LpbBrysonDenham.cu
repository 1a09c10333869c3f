// Kernel instantiation for the functor set LpbBrysonDenham (include/problems/bryson_denham.h).
#include "../../include/problems/bryson_denham.h"
#include "lpb_hessian.cuh"
LPB_DEFINE_FUNCTOR(LpbBrysonDenham)
