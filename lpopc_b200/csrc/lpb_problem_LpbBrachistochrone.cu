// Kernel instantiation for the functor set LpbBrachistochrone (include/problems/brachistochrone.h).
#include "../../include/problems/brachistochrone.h"
#include "lpb_hessian.cuh"
LPB_DEFINE_FUNCTOR(LpbBrachistochrone)
