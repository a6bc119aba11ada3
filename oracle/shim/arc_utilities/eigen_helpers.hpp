// TEST INFRASTRUCTURE.  simple_uncertainty_models.hpp includes arc_utilities/eigen_helpers.hpp but the actuator and sensor
// classes pinned by oracle/_ref/unc_ref use nothing from it: empty stand-in (see arc_helpers.hpp next to this file).
#ifndef FKS_SHIM_EIGEN_HELPERS_HPP
#define FKS_SHIM_EIGEN_HELPERS_HPP
#endif
