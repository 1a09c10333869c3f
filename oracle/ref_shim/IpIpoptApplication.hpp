// oracle/ref_shim/IpIpoptApplication.hpp -- TEST INFRASTRUCTURE ONLY.  Empty stand-in: the
// reference's LpNLPWrapper.cpp:11 includes the IPOPT header of this name but uses nothing from it.
#ifndef LPB_SHIM_IPAPP
#define LPB_SHIM_IPAPP
#include "IpTNLP.hpp"
#endif
