// oracle/ref_examples.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference's OWN example user-function classes (Lpopc/example/*/*.cpp, compiled unmodified by
// oracle/ref_build.mk with -Dmain="static lpopc_example_main", which turns each example's main() into an unused
// internal function) behind a factory, so that oracle/ref_driver.cpp can run the reference's transcription with
// the reference's user functions ("ref:<name>" functor names) -- Armadillo expressions and libm, not this
// repository's functor headers.  tests/test_reference_examples.py compares the functor headers
// (include/problems/{hypersensitive,bryson_denham,launch}.h), through the restatement and through the CUDA path,
// with these: a wrong constant, sign or term in a header shows up here.
//
//   HyperSensitiveFunction   example/hypersensitive/HyperSensitive.cpp:74-167
//   BrysonDenhamFunction     example/bryson-denham/BrysonDenham.cpp:95-167
//   LaunchFunction           example/launch/Launch.cpp:632-770, constants :11-74 (static initialisers) and
//                            :115-127,:148-153 (the part main() assigns; redone below because main() is not run)
#include "BrysonDenham.h"
#include "Launch.hpp"
#include "hypersensitive.h"

#include <memory>
#include <string>

// layout of the example's globals (Launch.cpp:23-34 and :50-66); the objects themselves live in Launch.o
struct struct_scales {
    double length, speed, time, acceleration, mass, force, area, volume, density, gravparam;
};
struct CONSTANTS_struct {
    double omega_matrix[9];
    double mu, cd, sa, rho0, H, Re, g0;
    double thrust_srb, thrust_first, thrust_second;
    double ISP_srb, ISP_first, ISP_second;
};
extern struct_scales scales;
extern CONSTANTS_struct CONSTANTS;

static void launch_main_constants()
{
    // Launch.cpp:89-91 burn times, :103-127 masses / thrusts / specific impulses, :148-153 the assignment
    const double bt_srb = 75.2 / scales.time, bt_first = 261.0 / scales.time, bt_second = 700.0 / scales.time;
    const double m_prop_srb = 17010 / scales.mass, m_prop_first = 95550 / scales.mass, m_prop_second = 16820 / scales.mass;
    const double thrust_srb = 628500 / scales.force, thrust_first = 1083100 / scales.force, thrust_second = 110094 / scales.force;
    const double mdot_srb = m_prop_srb / bt_srb, mdot_first = m_prop_first / bt_first, mdot_second = m_prop_second / bt_second;
    CONSTANTS.thrust_srb = thrust_srb;
    CONSTANTS.thrust_first = thrust_first;
    CONSTANTS.thrust_second = thrust_second;
    CONSTANTS.ISP_srb = thrust_srb / (CONSTANTS.g0 * mdot_srb);
    CONSTANTS.ISP_first = thrust_first / (CONSTANTS.g0 * mdot_first);
    CONSTANTS.ISP_second = thrust_second / (CONSTANTS.g0 * mdot_second);
}

std::shared_ptr<Lpopc::FunctionWrapper> lpopc_ref_example(const std::string& name)
{
    if (name == "hypersensitive") return std::shared_ptr<Lpopc::FunctionWrapper>(new HyperSensitiveFunction());
    if (name == "bryson_denham") return std::shared_ptr<Lpopc::FunctionWrapper>(new BrysonDenhamFunction());
    if (name == "launch") {
        launch_main_constants();
        return std::shared_ptr<Lpopc::FunctionWrapper>(new LaunchFunction());
    }
    return std::shared_ptr<Lpopc::FunctionWrapper>();
}

// the Launch example's constants in the order of LpbLaunch::Consts (include/problems/launch.h) plus the time scale
extern "C" int lpo_ref_launch_constants(double* out, int cap)
{
    launch_main_constants();
    const double v[15] = {CONSTANTS.omega_matrix[1], CONSTANTS.mu, CONSTANTS.cd, CONSTANTS.sa, CONSTANTS.rho0, CONSTANTS.H, CONSTANTS.Re,
                          CONSTANTS.g0, CONSTANTS.thrust_srb, CONSTANTS.thrust_first, CONSTANTS.thrust_second, CONSTANTS.ISP_srb,
                          CONSTANTS.ISP_first, CONSTANTS.ISP_second, scales.time};
    for (int i = 0; i < 15 && i < cap; ++i) out[i] = v[i];
    return 15;
}
