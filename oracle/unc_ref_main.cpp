// TEST INFRASTRUCTURE.  Driver around the REFERENCE's own simple_uncertainty_models.hpp, compiled from where it lies under
// /root/reference against the two-name stand-in of arc_utilities in oracle/shim (never copied into the repo).
// Reads "velocity_limit acceleration_limit proportional_noise minimum_noise percent_variance n" then n lines "control draw";
// prints per line the noiseless and the noisy control value (TruncatedNormalUncertainVelocityActuator::GetControlValue,
// unc.hpp:70-75 and :77-90) with %.17g, then one line with the standard deviation the reference hands its distribution
// (ClampValue(|percent_variance|, 0, 1), unc.hpp:61) and the truncation bounds.
#include <cassert>
#include <cmath>
#include <fast_kinematic_simulator/simple_uncertainty_models.hpp>

#include <cstdio>

struct ActuatorView : public simple_uncertainty_models::TruncatedNormalUncertainVelocityActuator {
    using simple_uncertainty_models::TruncatedNormalUncertainVelocityActuator::TruncatedNormalUncertainVelocityActuator;
    const arc_helpers::TruncatedNormalDistribution& Distribution() const { return noise_distribution_; }
};

int main() {
    double vl, al, pn, mn, pv;
    int n;
    if (std::scanf("%lf %lf %lf %lf %lf %d", &vl, &al, &pn, &mn, &pv, &n) != 6) return 1;
    const ActuatorView actuator(vl, al, pn, mn, pv);
    arc_helpers::TapeGenerator tape;
    std::vector<double> controls((size_t)n);
    tape.draws.resize((size_t)n);
    for (int i = 0; i < n; i++)
        if (std::scanf("%lf %lf", &controls[(size_t)i], &tape.draws[(size_t)i]) != 2) return 1;
    for (int i = 0; i < n; i++) {
        const double quiet = actuator.GetControlValue(controls[(size_t)i]);
        const double noisy = actuator.GetControlValue(controls[(size_t)i], tape);
        std::printf("%.17g %.17g\n", quiet, noisy);
    }
    const arc_helpers::TruncatedNormalDistribution& d = actuator.Distribution();
    std::printf("%.17g %.17g %.17g %.17g\n", d.mean, d.stddev, d.lower, d.upper);
    return 0;
}
