// CPU MODEL OF THE DEVICE SOLVER -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The product solves the stacked-Jacobian system (ComputeResolverCorrectionStepStackedJacobian,
// simple_particle_contact_simulator.hpp:1990-1998) with a warp-cooperative register QR
// (fast_kinematic_simulator_b200/csrc/fks_kernels.cu, qr_rolled): rows spread over the 32 lanes, column sums by
// butterfly trees, residual column norms recomputed every step instead of Eigen's downdating, tall systems
// folded 64 rows at a time by unpivoted reflections.  This file restates THAT algorithm on the CPU, with the
// floating-point contractions the device compiler makes (a * b + c -> fma) switchable, so that the question
// "does the device solver take the reference's rank decisions at the reference's rate?" can be answered on
// identical systems without a GPU (tests/test_oracle_qr_model.py), and so that whole trajectories can be run
// with it (FKS_ORACLE_QR_MODEL, see fks_oracle.cpp) to see what the arithmetic does to the aggregates.
//
// It follows the device code, not the reference: the Eigen-semantics solver is colpiv_qr_solve in
// fks_oracle.cpp.

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace fks_qr_model {

struct Options {
    bool fma;          // contract a * b + c the way nvcc does with -fmad=true
    bool fma_update;   // contraction of the reflector dot products / column updates only (the knob of the product)
    bool eigen_assoc;  // update as a -= (tau * v) * dot (Eigen's association) instead of a -= v * (tau * dot)
    bool sequential;   // sums over the rows in row order instead of per-lane partials + butterfly
    bool tail_beta;    // beta from c0^2 + (sum over the rows below the diagonal) instead of the residual norm of the pivot column
};

struct Result {
    int rank;
    int size;
    uint64_t order;    // nibble k: the column picked at step k
    double min_ratio;  // smallest pivot^2 / cut over the steps (>= 1: kept, < 1: cut)
};

namespace {

inline double mul_add(double a, double b, double c, bool fused) { return fused ? std::fma(a, b, c) : (a * b) + c; }

inline double sequential_sum(const double* v) {
    double s = 0.0;
    for (int l = 0; l < 32; l++) s += v[l];
    return s;
}
// sum of one value per lane over the 32 lanes by an xor butterfly (strides 16, 8, 4, 2, 1)
inline double butterfly(double* v) {
    for (int o = 16; o > 0; o >>= 1)
        for (int l = 0; l < 32; l++)
            if ((l & o) == 0) {
                const double s = v[l] + v[l ^ o];
                v[l] = s;
                v[l ^ o] = s;
            }
    return v[0];
}

// One pass of the register QR over `rows` <= 32 * R rows held in a[row][col] (row = lane + 32 * slot), NC unknowns,
// column NC = right-hand side.  reduce_only: unpivoted, all NC steps, leaves R in rows 0 .. NC-1.
struct Pass {
    int NC, R;
    std::vector<double> a;  // (32 R) x (NC + 1), row major
    double& at(int r, int c) { return a[(size_t)r * (NC + 1) + c]; }
};

int run_pass(Pass& P, int rows, int rows_thr, bool reduce_only, const Options& opt, double* x, Result* res,
             const uint64_t* forced_order, int forced_rank) {
    const int NC = P.NC, R = P.R;
    const int size = reduce_only ? NC : (rows < NC ? rows : NC);
    int nonzero_pivots = size;
    double threshold_helper = -1.0;
    int pos[16], order[16];
    double rdiag[16];
    for (int j = 0; j < 16; j++) pos[j] = j;
    double min_ratio = INFINITY;
    for (int k = 0; k < size; k++) {
        // tree 1: squared residual norms (rows >= k) of every column
        double nsq[16];
        for (int j = 0; j < NC; j++) {
            double lanes[32];
            for (int l = 0; l < 32; l++) {
                double v = (l >= k) ? P.at(l, j) * P.at(l, j) : 0.0;
                for (int sl = 1; sl < R; sl++) v = mul_add(P.at(l + 32 * sl, j), P.at(l + 32 * sl, j), v, opt.fma);
                lanes[l] = v;
            }
            nsq[j] = opt.sequential ? sequential_sum(lanes) : butterfly(lanes);
        }
        int p = k;
        double nsq_p = nsq[k];
        if (!reduce_only) {
            // arg-max over the columns not yet chosen, ties to the smallest position
            double best = -1.0;
            int best_pos = 1 << 30;
            p = -1;
            for (int j = 0; j < NC; j++) {
                if (pos[j] < k) continue;
                if (nsq[j] > best || (nsq[j] == best && pos[j] < best_pos)) {
                    best = nsq[j];
                    best_pos = pos[j];
                    p = j;
                }
            }
            if (forced_order) {
                p = (int)((*forced_order >> (4 * k)) & 0xFull);
                best = nsq[p];
                best_pos = pos[p];
            }
            const double big_sq = best;
            if (threshold_helper < 0.0) threshold_helper = (big_sq * (DBL_EPSILON * DBL_EPSILON)) / (double)rows_thr;
            const double cut = threshold_helper * (double)(rows_thr - k);
            if (nonzero_pivots == size && big_sq < cut) nonzero_pivots = k;
            if (cut > 0.0 && big_sq > 0.0) min_ratio = std::fmin(min_ratio, big_sq / cut);
            // the column that sat at position k takes the pivot's old position
            for (int j = 0; j < NC; j++)
                if (pos[j] == k && j != p) pos[j] = best_pos;
            pos[p] = k;
            nsq_p = big_sq;
        }
        if (forced_rank >= 0 && !reduce_only) nonzero_pivots = forced_rank < size ? forced_rank : size;
        order[k] = p;
        // makeHouseholderInPlace on the pivot column
        const int nrows = 32 * R;
        std::vector<double> v((size_t)nrows);
        for (int r = 0; r < nrows; r++) v[(size_t)r] = P.at(r, p);
        const double c0 = v[(size_t)k];
        bool has_tail = false;
        for (int r = k + 1; r < nrows; r++) has_tail = has_tail || v[(size_t)r] != 0.0;
        double tau = 0.0, beta = c0;
        if (has_tail) {
            if (opt.tail_beta) {
                double t = 0.0;
                for (int r = k + 1; r < nrows; r++) t += v[(size_t)r] * v[(size_t)r];
                nsq_p = c0 * c0 + t;
            }
            beta = std::sqrt(nsq_p);
            if (c0 >= 0.0) beta = -beta;
            const double inv_denom = 1.0 / (c0 - beta);
            for (int r = 0; r < nrows; r++) v[(size_t)r] = (r > k) ? v[(size_t)r] * inv_denom : 0.0;
        }
        const double inv_beta = 1.0 / beta;
        if (has_tail) tau = (beta - c0) * inv_beta;
        v[(size_t)k] = 1.0;
        rdiag[k] = inv_beta;
        if (reduce_only) P.at(k, p) = beta;
        const bool apply_b = reduce_only || nonzero_pivots > k;
        if (tau != 0.0) {
            for (int j = 0; j <= NC; j++) {
                const bool active = (j == NC) ? apply_b : (pos[j] > k);
                double lanes[32];
                for (int l = 0; l < 32; l++) {
                    double t = (l >= k) ? v[(size_t)l] * P.at(l, j) : 0.0;
                    for (int sl = 1; sl < R; sl++) t = mul_add(v[(size_t)(l + 32 * sl)], P.at(l + 32 * sl, j), t, opt.fma_update);
                    lanes[l] = t;
                }
                const double dtj = opt.sequential ? sequential_sum(lanes) : butterfly(lanes);
                if (!active) continue;
                if (opt.eigen_assoc) {
                    for (int r = k; r < nrows; r++) P.at(r, j) = mul_add(-(tau * v[(size_t)r]), dtj, P.at(r, j), opt.fma_update);
                    continue;
                }
                const double tmp = tau * dtj;
                for (int r = k; r < nrows; r++) P.at(r, j) = mul_add(-v[(size_t)r], tmp, P.at(r, j), opt.fma_update);
            }
        }
    }
    if (reduce_only) return 0;
    if (res) {
        res->rank = nonzero_pivots;
        res->size = size;
        res->order = 0;
        for (int k = 0; k < size; k++) res->order |= (uint64_t)order[k] << (4 * k);
        res->min_ratio = min_ratio;
    }
    for (int j = 0; j < NC; j++) x[j] = 0.0;
    std::vector<double> sres(32);
    for (int l = 0; l < 32; l++) sres[(size_t)l] = P.at(l, NC);
    for (int i = nonzero_pivots - 1; i >= 0; i--) {
        const int pc = order[i];
        const double yi = sres[(size_t)i] * rdiag[i];
        for (int l = 0; l < 32; l++) sres[(size_t)l] = mul_add(-P.at(l, pc), yi, sres[(size_t)l], opt.fma);
        x[pc] = yi;
    }
    return 0;
}

}  // namespace

// A: rows x cols column major, b: rows.  cols <= 15.  Returns the solution in x and the decisions in res.
void solve(const double* A, const double* b, int rows, int cols, const Options& opt, double* x, Result* res,
           const uint64_t* forced_order = nullptr, int forced_rank = -1) {
    // working copy with the right-hand side as column `cols`, row major
    std::vector<double> M((size_t)rows * (cols + 1));
    for (int r = 0; r < rows; r++) {
        for (int c = 0; c < cols; c++) M[(size_t)r * (cols + 1) + c] = A[(size_t)c * rows + r];
        M[(size_t)r * (cols + 1) + cols] = b[r];
    }
    int row0 = 0;
    // tall systems: fold 64 rows at a time into a cols x (cols + 1) triangle (unpivoted reflections)
    while (rows - row0 > 64) {
        Pass P;
        P.NC = cols;
        P.R = 2;
        P.a.assign((size_t)64 * (cols + 1), 0.0);
        std::memcpy(P.a.data(), M.data() + (size_t)row0 * (cols + 1), (size_t)64 * (cols + 1) * sizeof(double));
        run_pass(P, 64, 0, true, opt, nullptr, nullptr, nullptr, -1);
        for (int l = 0; l < cols; l++)
            for (int c = 0; c <= cols; c++) M[(size_t)(row0 + 64 - cols + l) * (cols + 1) + c] = (c < l) ? 0.0 : P.at(l, c);
        row0 += 64 - cols;
    }
    const int n = rows - row0;
    Pass P;
    P.NC = cols;
    P.R = n > 32 ? 2 : 1;
    P.a.assign((size_t)32 * P.R * (cols + 1), 0.0);
    std::memcpy(P.a.data(), M.data() + (size_t)row0 * (cols + 1), (size_t)n * (cols + 1) * sizeof(double));
    run_pass(P, n, rows, false, opt, x, res, forced_order, forced_rank);
}

}  // namespace fks_qr_model

extern "C" {
// opt_bits: bit 0 = fma everywhere, bit 1 = fma in the reflector dots / updates.  out4 = {rank, size, order, min_ratio}
void oracle_qr_device_model(const double* A_colmajor, const double* b, int rows, int cols, int opt_bits, double* x, double* out4) {
    fks_qr_model::Options opt;
    opt.fma = (opt_bits & 1) != 0;
    opt.fma_update = (opt_bits & 2) != 0;
    opt.eigen_assoc = (opt_bits & 8) != 0;
    opt.sequential = (opt_bits & 16) != 0;
    opt.tail_beta = (opt_bits & 32) != 0;
    fks_qr_model::Result r;
    fks_qr_model::solve(A_colmajor, b, rows, cols, opt, x, &r);
    if (out4) {
        out4[0] = (double)r.rank;
        out4[1] = (double)r.size;
        out4[2] = (double)r.order;
        out4[3] = r.min_ratio;
    }
}
// the same with the pivot order and the rank taken from the caller (what the device does with a decision tape)
void oracle_qr_device_model_forced(const double* A_colmajor, const double* b, int rows, int cols, int opt_bits, uint64_t order, int rank,
                                   double* x) {
    fks_qr_model::Options opt;
    opt.fma = (opt_bits & 1) != 0;
    opt.fma_update = (opt_bits & 2) != 0;
    opt.eigen_assoc = (opt_bits & 8) != 0;
    opt.sequential = (opt_bits & 16) != 0;
    opt.tail_beta = (opt_bits & 32) != 0;
    fks_qr_model::Result r;
    fks_qr_model::solve(A_colmajor, b, rows, cols, opt, x, &r, &order, rank);
}
}
