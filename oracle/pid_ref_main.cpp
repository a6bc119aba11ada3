// TEST INFRASTRUCTURE.  Driver around the REFERENCE's own simple_pid_controller.hpp (the only
// reference file that compiles stand-alone).  Built by oracle/Makefile into
// oracle/_ref/pid_ref from the header where it lies under /root/reference; never copied into the repo.
// Reads "kp ki kd iclamp n" then n lines "error timestep"; prints one feedback term per line (%.17g).
#include <fast_kinematic_simulator/simple_pid_controller.hpp>

#include <cstdio>

int main() {
    double kp, ki, kd, ic;
    int n;
    if (std::scanf("%lf %lf %lf %lf %d", &kp, &ki, &kd, &ic, &n) != 5) return 1;
    simple_pid_controller::SimplePIDController pid(kp, ki, kd, ic);
    for (int i = 0; i < n; i++) {
        double e, dt;
        if (std::scanf("%lf %lf", &e, &dt) != 2) return 1;
        std::printf("%.17g\n", pid.ComputeFeedbackTerm(e, dt));
    }
    return 0;
}
