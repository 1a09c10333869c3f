/* Compile-time registry of functor sets.  LPB_FOR_EACH_PROBLEM(X) expands X(Type)
 * once per registered problem; add a header + one line here to register a new
 * user problem (the device-functor equivalent of subclassing
 * Lpopc::FunctionWrapper, LpFunctionWrapper.h:50). */
#ifndef LPB_ALL_PROBLEMS_H
#define LPB_ALL_PROBLEMS_H
#include "hypersensitive.h"
#include "bryson_denham.h"
#include "launch.h"
#include "orbit_raising.h"
#include "brachistochrone.h"
#include "quadrotor.h"
#include "cartpole.h"
#include "synthetic20.h"
#include "two_stage.h"

#define LPB_FOR_EACH_PROBLEM(X) \
    X(LpbHypersensitive)        \
    X(LpbBrysonDenham)          \
    X(LpbLaunch)                \
    X(LpbOrbitRaising)          \
    X(LpbBrachistochrone)       \
    X(LpbQuadrotor)             \
    X(LpbCartpole)              \
    X(LpbSynthetic20)           \
    X(LpbTwoStage)
#endif
