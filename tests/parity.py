"""Shared helpers of the parity tests: run the CPU oracle with a recorded noise tape, replay the tape
on the GPU through the C ABI, compare.

Bar (BASELINE.json north_star): final configurations within 1e-9 relative (FP64); contact / failure
flags, controller-step, microstep and resolver-iteration counts bit-exact -- except for particles the
oracle marks SENSITIVE (some discrete decision came within tolerance of flipping: a voxel boundary, a
contact threshold, an angle at +-pi ...).  Those are enumerated, not asserted.

The oracle also records a DECISION TAPE (include/fksgpu.h, fks_noise_tape.decisions): rank and pivot order of
every stacked-Jacobian solve, plus the solution itself where a round-off pivot was kept (the reference divides by
round-off there; nothing reproduces it).  The GPU consumes it in injection mode, so those solver decisions are no
longer a reason to excuse a particle; the device reports where its own decision differed
(FKS_FLAG_DECISION_OVERRIDDEN) and whether the tape still fitted its trajectory (FKS_FLAG_DECISION_DESYNC).
`decision_cond_limit` additionally injects the solution of every solve whose condition estimate exceeds the limit
(ill-conditioned solves amplify last-bit differences of their INPUT by their condition number).
"""
import numpy as np

from fast_kinematic_simulator_b200 import capi
from oracle import oracle_binding as OB

RTOL = 1e-9
# flag bits with reference semantics (the device-only diagnostics NEAR_RANK_CUT / J_SPILLED are excluded)
SEMANTIC_FLAGS = (capi.FLAG_DID_CONTACT | capi.FLAG_RESOLVE_FAILED | capi.FLAG_ENDED_BY_FAILURE | capi.FLAG_ENDED_BY_NOCONTACT |
                  capi.FLAG_ENDED_BY_SHORTCUT | capi.FLAG_WOULD_ASSERT_MICROSTEP | capi.FLAG_WOULD_ASSERT_NORMAL |
                  capi.FLAG_WOULD_ASSERT_NAN | capi.FLAG_EMPTY_JACOBIAN | capi.FLAG_TAPE_EXHAUSTED)


def make_oracle(workload, solver_params=None, num_threads=0, seed=42):
    sp = solver_params if solver_params is not None else capi.default_solver_params()
    return OB.OracleSimulator(workload.environment().desc, workload.robot.to_c(), sp, 25.0, seed, num_threads)


def compare(gpu, ref, sens, rtol=RTOL, with_decisions=False):
    """gpu: SimulationResults; ref: oracle records; sens: oracle sensitivity mask.  Returns a report dict."""
    n = len(ref)
    if with_decisions:
        sens = sens & ~np.uint32(OB.SENS_COVERED_BY_DECISION_TAPE)
    g = gpu.records
    cfg_err = np.max(np.abs(g["cfg"] - ref["cfg"]) / np.maximum(1.0, np.abs(ref["cfg"])), axis=1)
    flags_ok = (g["flags"] & SEMANTIC_FLAGS) == (ref["flags"] & SEMANTIC_FLAGS)
    counts_ok = (g["n_microsteps"] == ref["n_microsteps"]) & (g["n_resolver_iters"] == ref["n_resolver_iters"]) & \
        (g["n_steps"] == ref["n_steps"])
    ok = flags_ok & counts_ok & (cfg_err <= rtol)
    insensitive = sens == 0
    return dict(
        n=n,
        n_insensitive=int(insensitive.sum()),
        n_match=int(ok.sum()),
        bad_insensitive=np.nonzero(~ok & insensitive)[0],
        bad_sensitive=np.nonzero(~ok & ~insensitive)[0],
        max_err_insensitive=float(cfg_err[insensitive].max()) if insensitive.any() else 0.0,
        max_err_matching=float(cfg_err[ok].max()) if ok.any() else 0.0,
        cfg_err=cfg_err,
        ok=ok,
        discrete_ok=flags_ok & counts_ok,
        n_overridden=int(((g["flags"] & capi.FLAG_DECISION_OVERRIDDEN) != 0).sum()),
        n_desync=int(((g["flags"] & capi.FLAG_DECISION_DESYNC) != 0).sum()),
    )


def describe(rep, sens):
    lines = ["particles %d, insensitive %d, matching %d; mismatching insensitive %d, mismatching sensitive (enumerated) %d; "
             "max rel err insensitive %.3g; decision tape: %d particles overridden, %d desynchronised" % (
                 rep["n"], rep["n_insensitive"], rep["n_match"], len(rep["bad_insensitive"]), len(rep["bad_sensitive"]),
                 rep["max_err_insensitive"], rep.get("n_overridden", 0), rep.get("n_desync", 0))]
    for i in rep["bad_sensitive"][:32]:
        bits = [nm for b, nm in enumerate(OB.SENS_NAMES) if (int(sens[i]) >> b) & 1]
        lines.append("  enumerated particle %d: sensitivity %s, rel err %.3g" % (i, "+".join(bits), rep["cfg_err"][i]))
    return "\n".join(lines)


def run_parity(workload, n, allow_contacts=True, solver_params=None, device=0, offset=0, decisions=True,
               decision_cond_limit=None):
    """oracle (mt19937 noise; noise and decision tapes recorded) -> GPU replay.
    Returns (report, gpu results, oracle records, sens)."""
    starts, targets = workload.subset(n, offset)
    orc = make_oracle(workload, solver_params)
    if decision_cond_limit is not None:
        OB.lib().oracle_set_decision_cond_limit(orc._h, float(decision_cond_limit))
    ref, tape, sens = OB.run_with_tape(orc, starts, targets, allow_contacts)
    ostats = orc.statistics()
    sim = workload.make_simulator(device=device, solver_params=solver_params)
    gpu = sim.forward_simulate_robots(starts, targets, allow_contacts, capi.NOISE_INJECTED, tape if decisions else tape[:2])
    gstats = sim.get_statistics()
    rep = compare(gpu, ref, sens, with_decisions=decisions)
    rep["decisions"] = OB.decision_records(tape, workload.robot.n_dof)
    rep["oracle_stats"] = {k: int(ostats[i]) for i, k in enumerate(capi.STAT_NAMES)}
    rep["gpu_stats"] = gstats
    sim.close()
    return rep, gpu, ref, sens
