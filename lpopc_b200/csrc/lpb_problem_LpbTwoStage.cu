// Kernel instantiation for the functor set LpbTwoStage (include/problems/two_stage.h).
#include "../../include/problems/two_stage.h"
#include "lpb_hessian.cuh"
LPB_DEFINE_FUNCTOR(LpbTwoStage)
