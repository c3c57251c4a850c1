// CPU restatement of the reference's per-period LinMPC path, used ONLY as the timed CPU baseline
// of bench.py (cpu_baseline / --impl reference) and checked against the numpy oracle in tests/.
// TEST / BENCH INFRASTRUCTURE -- never linked into libbmpc.so.
//
// The reference path is Julia + JuMP + the third-party OSQP C library (OSQP.jl compat "0.8",
// version unpinned, sources not under /root/reference); none of them exist in this image, so
// this file restates it ("restated OSQP-style ADMM -- OSQP binary unavailable"):
//   * initpred!        src/controller/execute.jl:247-277, INCLUDING the redundant M_Hp*Ẽ product
//                      recomputed every period (:263),
//   * linconstraint!   src/controller/transcription.jl:811-848 (full b vector, then b[i_b]),
//   * warm start       set_warmstart_mpc! transcription.jl:997-1007,
//   * the QP handed to the solver  linmpc.jl:323-339: rows A[i_b,:] z <= b[i_b] plus one row per
//                      finite variable bound (JuMP bridges variable bounds to affine rows),
//   * OSQP's published ADMM (Stellato et al. 2020, Algorithm 1) with the documented defaults
//     rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=eps_rel=1e-3, max_iter=4000, check_termination=25,
//     adaptive rho (tolerance 5, tested at the termination checks), 10 Ruiz equilibration passes,
//     primal warm start from Z̃s and dual warm start from the previous solve,
//   * getinput!        execute.jl:536-546.
// The linear system of each ADMM iteration is solved in the reduced form
//   (P + sigma I + rho A'A) x = sigma x_k - q + A'(rho z_k - y_k)
// (same iterates as OSQP's KKT form; cheaper than a dense (n+m) LDL' for these small dense
// problems, i.e. favourable to the CPU).  JuMP/MOI per-call overhead is not modelled (also
// favourable to the CPU).  Build: g++ -O3 -march=native -fopenmp -shared -fPIC.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Dims {
    int nu, ny, nx, Hp, Hc, neps, nY, nU, nz, n;
};

struct Solver {  // per instance, persistent across periods (like the JuMP model + OSQP workspace)
    int n = 0, m = 0;
    std::vector<double> P, A;        // scaled P (n x n, row-major), scaled A (m x n, row-major)
    std::vector<double> D, E;        // Ruiz scalings
    double c = 1.0;                  // cost scaling
    std::vector<double> L;           // Cholesky of P + sigma I + rho A'A
    double rho = 0.1;
    std::vector<double> x, z, y;     // scaled iterates (persist: warm start)
    std::vector<double> xt, zt, rhs, tmpm, tmpn, Ax, Px, Aty, q, l, u;
    long total_iters = 0, factorizations = 0;
};

constexpr double SIGMA = 1e-6, ALPHA = 1.6, RHO_MIN = 1e-6, RHO_MAX = 1e6;
double EPS_ABS = 1e-3, EPS_REL = 1e-3;  // OSQP defaults; cpuref_set_eps is for the convergence self-test only
int MAX_ITER = 4000;
constexpr int CHECK = 25, RUIZ = 10;

void factor(Solver& s) {
    const int n = s.n, m = s.m;
    std::vector<double>& L = s.L;
    L.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double a = s.P[(size_t)i * n + j];
            for (int r = 0; r < m; ++r) a += s.rho * s.A[(size_t)r * n + i] * s.A[(size_t)r * n + j];
            L[(size_t)i * n + j] = a;
        }
    for (int i = 0; i < n; ++i) L[(size_t)i * n + i] += SIGMA;
    for (int j = 0; j < n; ++j) {
        double d = L[(size_t)j * n + j];
        for (int p = 0; p < j; ++p) d -= L[(size_t)j * n + p] * L[(size_t)j * n + p];
        d = std::sqrt(std::max(d, 1e-300));
        L[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double a = L[(size_t)i * n + j];
            for (int p = 0; p < j; ++p) a -= L[(size_t)i * n + p] * L[(size_t)j * n + p];
            L[(size_t)i * n + j] = a / d;
        }
    }
    s.factorizations++;
}

void chol_solve(const Solver& s, double* b) {
    const int n = s.n;
    const double* L = s.L.data();
    for (int i = 0; i < n; ++i) {
        double a = b[i];
        for (int j = 0; j < i; ++j) a -= L[(size_t)i * n + j] * b[j];
        b[i] = a / L[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double a = b[i];
        for (int j = i + 1; j < n; ++j) a -= L[(size_t)j * n + i] * b[j];
        b[i] = a / L[(size_t)i * n + i];
    }
}

// setup: Ruiz equilibration of [P A'; A 0] (OSQP scaling = 10) + cost scaling, then factor
void setup(Solver& s, int n, int m, const double* P, const double* A) {
    s.n = n;
    s.m = m;
    s.P.assign(P, P + (size_t)n * n);
    s.A.assign(A, A + (size_t)m * n);
    s.D.assign(n, 1.0);
    s.E.assign(m, 1.0);
    std::vector<double> dn(n), em(m);
    for (int it = 0; it < RUIZ; ++it) {
        for (int j = 0; j < n; ++j) {
            double v = 0;
            for (int i = 0; i < n; ++i) v = std::max(v, std::fabs(s.P[(size_t)i * n + j]));
            for (int r = 0; r < m; ++r) v = std::max(v, std::fabs(s.A[(size_t)r * n + j]));
            dn[j] = 1.0 / std::sqrt(std::min(std::max(v, 1e-4), 1e4));
        }
        for (int r = 0; r < m; ++r) {
            double v = 0;
            for (int j = 0; j < n; ++j) v = std::max(v, std::fabs(s.A[(size_t)r * n + j]));
            em[r] = v < 1e-4 ? 1.0 : 1.0 / std::sqrt(std::min(v, 1e4));
        }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) s.P[(size_t)i * n + j] *= dn[i] * dn[j];
        for (int r = 0; r < m; ++r)
            for (int j = 0; j < n; ++j) s.A[(size_t)r * n + j] *= em[r] * dn[j];
        for (int j = 0; j < n; ++j) s.D[j] *= dn[j];
        for (int r = 0; r < m; ++r) s.E[r] *= em[r];
    }
    double cn = 0;  // mean column inf-norm of P
    for (int j = 0; j < n; ++j) {
        double v = 0;
        for (int i = 0; i < n; ++i) v = std::max(v, std::fabs(s.P[(size_t)i * n + j]));
        cn += v;
    }
    cn /= std::max(n, 1);
    s.c = 1.0 / std::min(std::max(cn, 1e-4), 1e4);
    for (auto& v : s.P) v *= s.c;
    s.rho = 0.1;
    s.x.assign(n, 0.0);
    s.z.assign(m, 0.0);
    s.y.assign(m, 0.0);
    s.xt.resize(n); s.zt.resize(m); s.rhs.resize(n); s.tmpm.resize(m); s.tmpn.resize(n);
    s.Ax.resize(m); s.Px.resize(n); s.Aty.resize(n); s.q.resize(n); s.l.resize(m); s.u.resize(m);
    factor(s);
}

// solve with (unscaled) q, l, u; xs = primal warm start (unscaled). Returns status 0 solved, 1 max iter.
int solve(Solver& s, const double* q, const double* l, const double* u, const double* xs, double* xout, int* iters) {
    const int n = s.n, m = s.m;
    for (int j = 0; j < n; ++j) {
        s.q[j] = s.c * s.D[j] * q[j];
        s.x[j] = xs[j] / s.D[j];
    }
    for (int r = 0; r < m; ++r) {
        s.l[r] = std::isfinite(l[r]) ? s.E[r] * l[r] : -1e30;
        s.u[r] = std::isfinite(u[r]) ? s.E[r] * u[r] : 1e30;
    }
    for (int r = 0; r < m; ++r) {  // z = A x (warm start)
        double a = 0;
        for (int j = 0; j < n; ++j) a += s.A[(size_t)r * n + j] * s.x[j];
        s.z[r] = std::min(std::max(a, s.l[r]), s.u[r]);
    }
    int status = 1, it = 0;
    for (it = 1; it <= MAX_ITER; ++it) {
        for (int r = 0; r < m; ++r) s.tmpm[r] = s.rho * s.z[r] - s.y[r];
        for (int j = 0; j < n; ++j) s.rhs[j] = SIGMA * s.x[j] - s.q[j];
        for (int r = 0; r < m; ++r) {
            const double w = s.tmpm[r];
            const double* ar = &s.A[(size_t)r * n];
            for (int j = 0; j < n; ++j) s.rhs[j] += ar[j] * w;
        }
        chol_solve(s, s.rhs.data());  // x tilde
        for (int r = 0; r < m; ++r) {
            double a = 0;
            const double* ar = &s.A[(size_t)r * n];
            for (int j = 0; j < n; ++j) a += ar[j] * s.rhs[j];
            s.zt[r] = a;  // z tilde = A x tilde
        }
        for (int j = 0; j < n; ++j) s.x[j] = ALPHA * s.rhs[j] + (1 - ALPHA) * s.x[j];
        for (int r = 0; r < m; ++r) {
            const double zr = ALPHA * s.zt[r] + (1 - ALPHA) * s.z[r];
            const double zn = std::min(std::max(zr + s.y[r] / s.rho, s.l[r]), s.u[r]);
            s.y[r] += s.rho * (zr - zn);
            s.z[r] = zn;
        }
        if (it % CHECK == 0 || it == MAX_ITER) {
            // unscaled residuals (scaled_termination = false)
            double rp = 0, rdn = 0, nAx = 0, nz = 0, nPx = 0, nAty = 0, nq = 0;
            for (int r = 0; r < m; ++r) {
                double a = 0;
                const double* ar = &s.A[(size_t)r * n];
                for (int j = 0; j < n; ++j) a += ar[j] * s.x[j];
                s.Ax[r] = a;
                rp = std::max(rp, std::fabs(a - s.z[r]) / s.E[r]);
                nAx = std::max(nAx, std::fabs(a) / s.E[r]);
                nz = std::max(nz, std::fabs(s.z[r]) / s.E[r]);
            }
            std::fill(s.Aty.begin(), s.Aty.end(), 0.0);
            for (int r = 0; r < m; ++r) {
                const double* ar = &s.A[(size_t)r * n];
                for (int j = 0; j < n; ++j) s.Aty[j] += ar[j] * s.y[r];
            }
            for (int i = 0; i < n; ++i) {
                double a = 0;
                for (int j = 0; j < n; ++j) a += s.P[(size_t)i * n + j] * s.x[j];
                s.Px[i] = a;
                const double sc = 1.0 / (s.c * s.D[i]);
                rdn = std::max(rdn, std::fabs(a + s.q[i] + s.Aty[i]) * sc);
                nPx = std::max(nPx, std::fabs(a) * sc);
                nAty = std::max(nAty, std::fabs(s.Aty[i]) * sc);
                nq = std::max(nq, std::fabs(s.q[i]) * sc);
            }
            const double ep = EPS_ABS + EPS_REL * std::max(nAx, nz);
            const double ed = EPS_ABS + EPS_REL * std::max(std::max(nPx, nAty), nq);
            if (rp <= ep && rdn <= ed) {
                status = 0;
                break;
            }
            // adaptive rho (scaled residuals, OSQP compute_rho_estimate)
            double srp = 0, srd = 0, sAx = 0, sz = 0, sPx = 0, sAty = 0, sq = 0;
            for (int r = 0; r < m; ++r) {
                srp = std::max(srp, std::fabs(s.Ax[r] - s.z[r]));
                sAx = std::max(sAx, std::fabs(s.Ax[r]));
                sz = std::max(sz, std::fabs(s.z[r]));
            }
            for (int i = 0; i < n; ++i) {
                srd = std::max(srd, std::fabs(s.Px[i] + s.q[i] + s.Aty[i]));
                sPx = std::max(sPx, std::fabs(s.Px[i]));
                sAty = std::max(sAty, std::fabs(s.Aty[i]));
                sq = std::max(sq, std::fabs(s.q[i]));
            }
            const double pn = srp / (std::max(sAx, sz) + 1e-10), dn = srd / (std::max(std::max(sPx, sAty), sq) + 1e-10);
            double rn = s.rho * std::sqrt(pn / (dn + 1e-10));
            rn = std::min(std::max(rn, RHO_MIN), RHO_MAX);
            if (rn > 5 * s.rho || rn < s.rho / 5) {
                s.rho = rn;
                factor(s);
            }
        }
    }
    it = std::min(it, MAX_ITER);
    for (int j = 0; j < n; ++j) xout[j] = s.D[j] * s.x[j];
    *iters = it;
    s.total_iters += it;
    return status;
}

}  // namespace

extern "C" {

// Runs `steps` replayed control periods for instances [0, N) on `threads` OpenMP threads.
// Inputs (instance-major, column-major matrices as in include/bmpc.h):
//   E (N x nY x nz), K (N x nY x nx), V (N x nY x nu), B (N x nY), Ht (N x n x n), Mdiag (N x nY)
//   Acon (m_all x n, row-major, SHARED structure incl. softness columns; rows: Umin Umax DUmin DUmax Ymin Ymax xmin xmax
//         where the Y blocks are per-instance -E / +E and the x̂ blocks are omitted -> nx rows unsupported here)
//   bounds U0min.. (N x len), i_b mask (m_all), Zmin/Zmax (N x n)
//   per step t: xhat0[t] (N x nx), lastu0[t] (N x nu), ry[t] (N x ny)
// Outputs: Zout (steps x N x n), uout (steps x N x nu), iters (steps x N), seconds (wall).
int cpuref_linmpc_run(int N, int steps, int threads, int nu, int ny, int nx, int Hp, int Hc, int neps,
                      const int* nb, const double* E, const double* K, const double* V, const double* B,
                      const double* Ht, const double* Mdiag, const double* U0min, const double* U0max,
                      const double* DUmin, const double* DUmax, const double* Y0min, const double* Y0max,
                      const double* C_umin, const double* C_umax, const double* C_dumin, const double* C_dumax,
                      const double* C_ymin, const double* C_ymax, const double* yop, const double* xhat0,
                      const double* lastu0, const double* ry, double* Zout, double* uout, int32_t* iters_out,
                      int32_t* status_out, double* seconds, int64_t* total_admm_iters) {
    Dims d{nu, ny, nx, Hp, Hc, neps, ny * Hp, nu * Hp, nu * Hc, nu * Hc + neps};
    const int nY = d.nY, nU = d.nU, nz = d.nz, n = d.n;
    std::vector<int> blk(Hp);
    {
        int t = 0;
        for (int l = 0; l < Hc; ++l)
            for (int k = 0; k < nb[l]; ++k) blk[t++] = l;
    }
    const int m_all = 2 * nU + 2 * nz + 2 * nY;
    std::vector<Solver> solvers(N);
    std::vector<std::vector<int>> rows_of(N);  // selected rows (i_b) per instance + bound rows
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    // ---- setup (not timed): build A per instance exactly as init_matconstraint_mpc + relax* ----
    std::vector<std::vector<double>> Afull(N);
    std::vector<std::vector<unsigned char>> ib(N);
    std::vector<std::vector<double>> Zmin(N), Zmax(N);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        std::vector<double>& A = Afull[i];
        A.assign((size_t)m_all * n, 0.0);
        auto at = [&](int r, int c) -> double& { return A[(size_t)r * n + c]; };
        const double* Ei = E + (size_t)i * nY * nz;
        int r0 = 0;
        for (int t = 0; t < Hp; ++t)
            for (int ch = 0; ch < nu; ++ch) {
                const int r = t * nu + ch;
                for (int l = 0; l <= blk[t]; ++l) {
                    at(r0 + r, l * nu + ch) = -1.0;
                    at(r0 + nU + r, l * nu + ch) = 1.0;
                }
                if (neps) {
                    at(r0 + r, nz) = -C_umin[r];
                    at(r0 + nU + r, nz) = -C_umax[r];
                }
            }
        r0 += 2 * nU;
        for (int k = 0; k < nz; ++k) {
            at(r0 + k, k) = -1.0;
            at(r0 + nz + k, k) = 1.0;
            if (neps) {
                at(r0 + k, nz) = -C_dumin[k];
                at(r0 + nz + k, nz) = -C_dumax[k];
            }
        }
        r0 += 2 * nz;
        for (int t = 0; t < nY; ++t) {
            for (int j = 0; j < nz; ++j) {
                at(r0 + t, j) = -Ei[t + (size_t)nY * j];
                at(r0 + nY + t, j) = Ei[t + (size_t)nY * j];
            }
            if (neps) {
                at(r0 + t, nz) = -C_ymin[t];
                at(r0 + nY + t, nz) = -C_ymax[t];
            }
        }
        // box constraints + i_b (init_boxconstraint_mpc, deleteΔU_lincon!)
        Zmin[i].assign(n, -INFINITY);
        Zmax[i].assign(n, INFINITY);
        if (neps) Zmin[i][nz] = 0.0;
        ib[i].assign(m_all, 0);
        for (int k = 0; k < nU; ++k) {
            ib[i][k] = std::isfinite(U0min[(size_t)i * nU + k]);
            ib[i][nU + k] = std::isfinite(U0max[(size_t)i * nU + k]);
        }
        for (int k = 0; k < nz; ++k) {
            const double lo = DUmin[(size_t)i * nz + k], hi = DUmax[(size_t)i * nz + k];
            const bool hard_lo = !neps || C_dumin[k] == 0.0, hard_hi = !neps || C_dumax[k] == 0.0;
            if (hard_lo) Zmin[i][k] = lo; else ib[i][2 * nU + k] = std::isfinite(lo);
            if (hard_hi) Zmax[i][k] = hi; else ib[i][2 * nU + nz + k] = std::isfinite(hi);
        }
        for (int t = 0; t < nY; ++t) {
            ib[i][2 * nU + 2 * nz + t] = std::isfinite(Y0min[(size_t)i * nY + t]);
            ib[i][2 * nU + 2 * nz + nY + t] = std::isfinite(Y0max[(size_t)i * nY + t]);
        }
        // OSQP problem: selected rows, then one row per finite variable bound
        std::vector<double> Aq;
        int m = 0;
        for (int r = 0; r < m_all; ++r)
            if (ib[i][r]) {
                Aq.insert(Aq.end(), A.begin() + (size_t)r * n, A.begin() + (size_t)(r + 1) * n);
                rows_of[i].push_back(r);
                ++m;
            }
        for (int k = 0; k < n; ++k)
            for (int side = 0; side < 2; ++side) {
                const double v = side ? Zmax[i][k] : Zmin[i][k];
                if (std::isfinite(v)) {
                    std::vector<double> row(n, 0.0);
                    row[k] = 1.0;
                    Aq.insert(Aq.end(), row.begin(), row.end());
                    rows_of[i].push_back(-(2 * k + side) - 1);
                    ++m;
                }
            }
        std::vector<double> P((size_t)n * n);
        const double* H = Ht + (size_t)i * n * n;
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < n; ++b) P[(size_t)a * n + b] = a >= b ? H[a + (size_t)n * b] : H[b + (size_t)n * a];
        setup(solvers[i], n, m, P.data(), Aq.data());
    }
    // ---- timed region: the per-period path ----
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel
    {
        std::vector<double> F(nY), Cy(nY), MEt((size_t)nY * n), q(n), b(m_all), Zs(n), Zprev(n, 0.0), lvec, uvec, zsol(n);
#pragma omp for schedule(dynamic, 4)
        for (int i = 0; i < N; ++i) {
            Solver& s = solvers[i];
            lvec.assign(s.m, -INFINITY);
            uvec.assign(s.m, INFINITY);
            std::fill(Zprev.begin(), Zprev.end(), 0.0);
            const double* Ei = E + (size_t)i * nY * nz;
            const double* Ki = K + (size_t)i * nY * nx;
            const double* Vi = V + (size_t)i * nY * nu;
            const double* Bi = B + (size_t)i * nY;
            const double* Mi = Mdiag + (size_t)i * nY;
            for (int t = 0; t < steps; ++t) {
                const double* xh = xhat0 + ((size_t)t * N + i) * nx;
                const double* lu = lastu0 + ((size_t)t * N + i) * nu;
                const double* r = ry + ((size_t)t * N + i) * ny;
                // initpred!
                for (int k = 0; k < nY; ++k) {
                    double f = Bi[k];
                    for (int j = 0; j < nx; ++j) f += Ki[k + (size_t)nY * j] * xh[j];
                    for (int j = 0; j < nu; ++j) f += Vi[k + (size_t)nY * j] * lu[j];
                    F[k] = f;
                    Cy[k] = f + yop[(size_t)i * ny + k % ny] - r[k % ny];
                }
                for (int k = 0; k < nY; ++k)  // M_Hp*Ẽ, recomputed every period as in the reference (:263)
                    for (int j = 0; j < n; ++j) MEt[(size_t)k * n + j] = j < nz ? Mi[k] * Ei[k + (size_t)nY * j] : 0.0;
                for (int j = 0; j < n; ++j) q[j] = 0.0;
                for (int k = 0; k < nY; ++k)
                    for (int j = 0; j < n; ++j) q[j] += MEt[(size_t)k * n + j] * Cy[k];
                for (int j = 0; j < n; ++j) q[j] *= 2.0;
                // linconstraint!
                for (int k = 0; k < nU; ++k) {
                    b[k] = -U0min[(size_t)i * nU + k] + lu[k % nu];
                    b[nU + k] = U0max[(size_t)i * nU + k] - lu[k % nu];
                }
                for (int k = 0; k < nz; ++k) {
                    b[2 * nU + k] = -DUmin[(size_t)i * nz + k];
                    b[2 * nU + nz + k] = DUmax[(size_t)i * nz + k];
                }
                for (int k = 0; k < nY; ++k) {
                    b[2 * nU + 2 * nz + k] = -Y0min[(size_t)i * nY + k] + F[k];
                    b[2 * nU + 2 * nz + nY + k] = Y0max[(size_t)i * nY + k] - F[k];
                }
                for (int rr = 0; rr < s.m; ++rr) {
                    const int src = rows_of[i][rr];
                    if (src >= 0) {
                        uvec[rr] = b[src];
                    } else {
                        const int code = -src - 1, k = code / 2, side = code % 2;
                        if (side) uvec[rr] = Zmax[i][k]; else lvec[rr] = Zmin[i][k];
                    }
                }
                // warm start
                for (int j = 0; j < nz; ++j) Zs[j] = j + nu < nz ? Zprev[j + nu] : 0.0;
                if (neps) Zs[nz] = Zprev[nz];
                int it = 0;
                const int st = solve(s, q.data(), lvec.data(), uvec.data(), Zs.data(), zsol.data(), &it);
                Zprev = zsol;
                double* Zo = Zout + ((size_t)t * N + i) * n;
                for (int j = 0; j < n; ++j) Zo[j] = zsol[j];
                for (int j = 0; j < nu; ++j) uout[((size_t)t * N + i) * nu + j] = zsol[j] + lu[j];
                iters_out[(size_t)t * N + i] = it;
                status_out[(size_t)t * N + i] = st;
            }
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    *seconds = std::chrono::duration<double>(t1 - t0).count();
    int64_t tot = 0;
    for (auto& s : solvers) tot += s.total_iters;
    *total_admm_iters = tot;
    return 0;
}

void cpuref_set_eps(double eps, int max_iter) {
    EPS_ABS = EPS_REL = eps;
    MAX_ITER = max_iter;
}

int cpuref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
}
