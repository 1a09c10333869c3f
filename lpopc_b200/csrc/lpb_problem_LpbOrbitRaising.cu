// Kernel instantiation for the functor set LpbOrbitRaising (include/problems/orbit_raising.h).
#include "../../include/problems/orbit_raising.h"
#include "lpb_hessian.cuh"
LPB_DEFINE_FUNCTOR(LpbOrbitRaising)
