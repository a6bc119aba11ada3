// TEST INFRASTRUCTURE.  Clean-room stand-in for the two names of arc_utilities/arc_helpers.hpp that the reference's
// simple_uncertainty_models.hpp uses (arc_utilities is not vendored by the reference and absent from this image), so that the
// reference's own actuator code compiles where it lies (oracle/Makefile -> oracle/_ref/unc_ref):
//   ClampValue                    call sites unc.hpp:61,74 -- min(max(value, low), high)
//   TruncatedNormalDistribution   call sites unc.hpp:25,53,61 -- constructed (mean, stddev, lower, upper), sampled as dist(rng).
// The stand-in distribution does NOT sample: it hands out the next value of a tape the "generator" carries, so that the
// reference's arithmetic around the draw (unc.hpp:77-90) is exercised with injected draws -- the same protocol as the GPU's
// FKS_NOISE_INJECTED mode.
#ifndef FKS_SHIM_ARC_HELPERS_HPP
#define FKS_SHIM_ARC_HELPERS_HPP
// (the real header pulls these standard headers in; the reference relies on that for std::shared_ptr, assert, std::abs)
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <map>
#include <memory>
#include <vector>

namespace arc_helpers {
template <typename T>
inline T ClampValue(const T& val, const T& min, const T& max) {
    return val < min ? min : (val > max ? max : val);
}
struct TapeGenerator {
    std::vector<double> draws;
    size_t pos = 0;
};
class TruncatedNormalDistribution {
public:
    double mean, stddev, lower, upper;
    TruncatedNormalDistribution(double m, double s, double lo, double hi) : mean(m), stddev(s), lower(lo), upper(hi) {}
    double operator()(TapeGenerator& g) { return g.draws[g.pos++]; }
};
}  // namespace arc_helpers
#endif
