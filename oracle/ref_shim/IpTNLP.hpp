// oracle/ref_shim/IpTNLP.hpp -- TEST INFRASTRUCTURE ONLY.  Declarations-only stand-in for the
// IPOPT header of the same name so that the reference's LpopcIpopt.h (included by
// LpNLPWrapper.cpp:12, never used by it) parses.  IPOPT is an un-vendored dependency of the
// reference (Lpopc/CMakeLists.txt:6,43) and is not in this image.  Nothing here is executed.
#ifndef LPB_SHIM_IPTNLP
#define LPB_SHIM_IPTNLP
namespace Ipopt {
typedef int Index;
typedef double Number;
enum SolverReturn { SUCCESS = 0 };
class IpoptData;
class IpoptCalculatedQuantities;
class TNLP {
public:
    enum IndexStyleEnum { C_STYLE = 0, FORTRAN_STYLE = 1 };
    virtual ~TNLP() {}
};
} // namespace Ipopt
#endif
