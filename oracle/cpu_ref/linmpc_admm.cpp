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
    std::vector<double> xt, zt, rhs, tmpm, tmpn, Ax, Px, Aty, q, l, u, dy;
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
    s.Ax.resize(m); s.Px.resize(n); s.Aty.resize(n); s.q.resize(n); s.l.resize(m); s.u.resize(m); s.dy.assign(m, 0.0);
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
            s.dy[r] = s.rho * (zr - zn);
            s.y[r] += s.dy[r];
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
            // primal infeasibility certificate (OSQP is_primal_infeasible, eps_prim_inf = 1e-4): dy with
            // ||A' dy|| <= eps ||dy||  and  u'(dy)+ + l'(dy)- <= -eps ||dy||   (unscaled norms)
            {
                const double EPS_PINF = 1e-4;
                double ndy = 0;
                for (int r = 0; r < m; ++r) ndy = std::max(ndy, std::fabs(s.E[r] * s.dy[r]));
                if (ndy > EPS_PINF * EPS_PINF) {
                    double lhs = 0;
                    bool cert = true;
                    for (int r = 0; r < m && cert; ++r) {
                        const double d = s.dy[r];
                        if (d > 0) {
                            if (s.u[r] >= 1e29) cert = d * s.E[r] <= EPS_PINF * ndy; else lhs += s.u[r] * d;
                        } else if (d < 0) {
                            if (s.l[r] <= -1e29) cert = -d * s.E[r] <= EPS_PINF * ndy; else lhs += s.l[r] * d;
                        }
                    }
                    if (cert && lhs < -EPS_PINF * ndy) {
                        double nat = 0;
                        for (int j = 0; j < n; ++j) {
                            double a = 0;
                            for (int r = 0; r < m; ++r) a += s.A[(size_t)r * n + j] * s.dy[r];
                            nat = std::max(nat, std::fabs(a) / s.D[j]);
                        }
                        if (nat <= EPS_PINF * ndy) {
                            status = 3;  // PRIMAL_INFEASIBLE
                            break;
                        }
                    }
                }
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

// One controller instance: constant data, the persistent solver workspace (like the JuMP model + OSQP workspace of
// the reference) and the per-period path.
struct Inst {
    Dims d;
    const double *E, *K, *V, *B, *M, *U0min, *U0max, *DUmin, *DUmax, *Y0min, *Y0max, *yop;
    Solver s;
    std::vector<int> rows_of;  // selected rows (i_b) + bound rows
    std::vector<double> Zmin, Zmax, Zprev;
    // scratch
    std::vector<double> F, Cy, MEt, q, b, Zs, lvec, uvec, zsol;

    // setup (not timed): A exactly as init_matconstraint_mpc + relax*, i_b, box constraints, OSQP setup
    void setup(const std::vector<int>& blk, const double* Ht, const double* C_umin, const double* C_umax,
               const double* C_dumin, const double* C_dumax, const double* C_ymin, const double* C_ymax) {
        const int nY = d.nY, nU = d.nU, nz = d.nz, n = d.n, nu = d.nu, Hp = d.Hp, neps = d.neps;
        const int m_all = 2 * nU + 2 * nz + 2 * nY;
        std::vector<double> A((size_t)m_all * n, 0.0);
        auto at = [&](int r, int c) -> double& { return A[(size_t)r * n + c]; };
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
                at(r0 + t, j) = -E[t + (size_t)nY * j];
                at(r0 + nY + t, j) = E[t + (size_t)nY * j];
            }
            if (neps) {
                at(r0 + t, nz) = -C_ymin[t];
                at(r0 + nY + t, nz) = -C_ymax[t];
            }
        }
        // box constraints + i_b (init_boxconstraint_mpc, deleteΔU_lincon!)
        Zmin.assign(n, -INFINITY);
        Zmax.assign(n, INFINITY);
        if (neps) Zmin[nz] = 0.0;
        std::vector<unsigned char> ib(m_all, 0);
        for (int k = 0; k < nU; ++k) {
            ib[k] = std::isfinite(U0min[k]);
            ib[nU + k] = std::isfinite(U0max[k]);
        }
        for (int k = 0; k < nz; ++k) {
            const double lo = DUmin[k], hi = DUmax[k];
            const bool hard_lo = !neps || C_dumin[k] == 0.0, hard_hi = !neps || C_dumax[k] == 0.0;
            if (hard_lo) Zmin[k] = lo; else ib[2 * nU + k] = std::isfinite(lo);
            if (hard_hi) Zmax[k] = hi; else ib[2 * nU + nz + k] = std::isfinite(hi);
        }
        for (int t = 0; t < nY; ++t) {
            ib[2 * nU + 2 * nz + t] = std::isfinite(Y0min[t]);
            ib[2 * nU + 2 * nz + nY + t] = std::isfinite(Y0max[t]);
        }
        // OSQP problem: selected rows, then one row per finite variable bound
        std::vector<double> Aq;
        int m = 0;
        for (int r = 0; r < m_all; ++r)
            if (ib[r]) {
                Aq.insert(Aq.end(), A.begin() + (size_t)r * n, A.begin() + (size_t)(r + 1) * n);
                rows_of.push_back(r);
                ++m;
            }
        for (int k = 0; k < n; ++k)
            for (int side = 0; side < 2; ++side) {
                const double v = side ? Zmax[k] : Zmin[k];
                if (std::isfinite(v)) {
                    std::vector<double> row(n, 0.0);
                    row[k] = 1.0;
                    Aq.insert(Aq.end(), row.begin(), row.end());
                    rows_of.push_back(-(2 * k + side) - 1);
                    ++m;
                }
            }
        std::vector<double> P((size_t)n * n);
        for (int a = 0; a < n; ++a)
            for (int b2 = 0; b2 < n; ++b2) P[(size_t)a * n + b2] = a >= b2 ? Ht[a + (size_t)n * b2] : Ht[b2 + (size_t)n * a];
        ::setup(s, n, m, P.data(), Aq.data());
        F.resize(nY); Cy.resize(nY); MEt.resize((size_t)nY * n); q.resize(n); b.resize(m_all); Zs.resize(n);
        Zprev.assign(n, 0.0); zsol.resize(n);
        lvec.assign(s.m, -INFINITY);
        uvec.assign(s.m, INFINITY);
    }

    // the per-period path: initpred!, linconstraint!, warm start, solve, getinput!
    int period(const double* xh, const double* lu, const double* r, double* Zo, double* uo, int* it_out) {
        const int nY = d.nY, nU = d.nU, nz = d.nz, n = d.n, nu = d.nu, ny = d.ny, nx = d.nx, neps = d.neps;
        for (int k = 0; k < nY; ++k) {
            double f = B[k];
            for (int j = 0; j < nx; ++j) f += K[k + (size_t)nY * j] * xh[j];
            for (int j = 0; j < nu; ++j) f += V[k + (size_t)nY * j] * lu[j];
            F[k] = f;
            Cy[k] = f + yop[k % ny] - r[k % ny];
        }
        for (int k = 0; k < nY; ++k)  // M_Hp*Ẽ, recomputed every period as in the reference (:263)
            for (int j = 0; j < n; ++j) MEt[(size_t)k * n + j] = j < nz ? M[k] * E[k + (size_t)nY * j] : 0.0;
        for (int j = 0; j < n; ++j) q[j] = 0.0;
        for (int k = 0; k < nY; ++k)
            for (int j = 0; j < n; ++j) q[j] += MEt[(size_t)k * n + j] * Cy[k];
        for (int j = 0; j < n; ++j) q[j] *= 2.0;
        // linconstraint!
        for (int k = 0; k < nU; ++k) {
            b[k] = -U0min[k] + lu[k % nu];
            b[nU + k] = U0max[k] - lu[k % nu];
        }
        for (int k = 0; k < nz; ++k) {
            b[2 * nU + k] = -DUmin[k];
            b[2 * nU + nz + k] = DUmax[k];
        }
        for (int k = 0; k < nY; ++k) {
            b[2 * nU + 2 * nz + k] = -Y0min[k] + F[k];
            b[2 * nU + 2 * nz + nY + k] = Y0max[k] - F[k];
        }
        for (int rr = 0; rr < s.m; ++rr) {
            const int src = rows_of[rr];
            if (src >= 0) {
                uvec[rr] = b[src];
            } else {
                const int code = -src - 1, k = code / 2, side = code % 2;
                if (side) uvec[rr] = Zmax[k]; else lvec[rr] = Zmin[k];
            }
        }
        // warm start
        for (int j = 0; j < nz; ++j) Zs[j] = j + nu < nz ? Zprev[j + nu] : 0.0;
        if (neps) Zs[nz] = Zprev[nz];
        int it = 0;
        const int st = solve(s, q.data(), lvec.data(), uvec.data(), Zs.data(), zsol.data(), &it);
        Zprev = zsol;
        if (Zo) for (int j = 0; j < n; ++j) Zo[j] = zsol[j];
        for (int j = 0; j < nu; ++j) uo[j] = zsol[j] + lu[j];
        *it_out = it;
        return st;
    }
};

std::vector<int> block_of_step(int Hp, int Hc, const int* nb) {
    std::vector<int> blk(Hp);
    int t = 0;
    for (int l = 0; l < Hc; ++l)
        for (int k = 0; k < nb[l]; ++k) blk[t++] = l;
    return blk;
}

void bind(Inst& I, const Dims& d, int i, const double* E, const double* K, const double* V, const double* B,
          const double* Mdiag, const double* U0min, const double* U0max, const double* DUmin, const double* DUmax,
          const double* Y0min, const double* Y0max, const double* yop) {
    I.d = d;
    I.E = E + (size_t)i * d.nY * d.nz;
    I.K = K + (size_t)i * d.nY * d.nx;
    I.V = V + (size_t)i * d.nY * d.nu;
    I.B = B + (size_t)i * d.nY;
    I.M = Mdiag + (size_t)i * d.nY;
    I.U0min = U0min + (size_t)i * d.nU;
    I.U0max = U0max + (size_t)i * d.nU;
    I.DUmin = DUmin + (size_t)i * d.nz;
    I.DUmax = DUmax + (size_t)i * d.nz;
    I.Y0min = Y0min + (size_t)i * d.nY;
    I.Y0max = Y0max + (size_t)i * d.nY;
    I.yop = yop + (size_t)i * d.ny;
}

}  // namespace

extern "C" {

// Runs `steps` replayed control periods for instances [0, N) on `threads` OpenMP threads; the first `warm` of them untimed.
// Inputs (instance-major, column-major matrices as in include/bmpc.h):
//   E (N x nY x nz), K (N x nY x nx), V (N x nY x nu), B (N x nY), Ht (N x n x n), Mdiag (N x nY),
//   bounds U0min .. Y0max (N x len), softness C_* (len, shared), yop (N x ny);
//   per step t: xhat0[t] (N x nx), lastu0[t] (N x nu), ry[t] (N x ny)
// Outputs: Zout (steps x N x n), uout (steps x N x nu), iters (steps x N), seconds (wall time of the per-period path).
int cpuref_linmpc_run(int N, int steps, int warm, int threads, int nu, int ny, int nx, int Hp, int Hc, int neps,
                      const int* nb, const double* E, const double* K, const double* V, const double* B,
                      const double* Ht, const double* Mdiag, const double* U0min, const double* U0max,
                      const double* DUmin, const double* DUmax, const double* Y0min, const double* Y0max,
                      const double* C_umin, const double* C_umax, const double* C_dumin, const double* C_dumax,
                      const double* C_ymin, const double* C_ymax, const double* yop, const double* xhat0,
                      const double* lastu0, const double* ry, double* Zout, double* uout, int32_t* iters_out,
                      int32_t* status_out, double* seconds, int64_t* total_admm_iters) {
    Dims d{nu, ny, nx, Hp, Hc, neps, ny * Hp, nu * Hp, nu * Hc, nu * Hc + neps};
    const int n = d.n;
    const std::vector<int> blk = block_of_step(Hp, Hc, nb);
    std::vector<Inst> inst(N);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        bind(inst[i], d, i, E, K, V, B, Mdiag, U0min, U0max, DUmin, DUmax, Y0min, Y0max, yop);
        inst[i].setup(blk, Ht + (size_t)i * n * n, C_umin, C_umax, C_dumin, C_dumax, C_ymin, C_ymax);
    }
    // ---- untimed: the first `warm` periods (they leave the solver workspaces warm, as in a running closed loop) ----
    warm = std::max(0, std::min(warm, steps));
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < N; ++i) {
        Inst& I = inst[i];
        for (int t = 0; t < warm; ++t) {
            const size_t o = (size_t)t * N + i;
            int it = 0;
            const int st = I.period(xhat0 + o * nx, lastu0 + o * nu, ry + o * ny, Zout + o * n, uout + o * nu, &it);
            iters_out[o] = it;
            status_out[o] = st;
        }
    }
    // ---- timed region: the per-period path ----
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < N; ++i) {
        Inst& I = inst[i];
        for (int t = warm; t < steps; ++t) {
            const size_t o = (size_t)t * N + i;
            int it = 0;
            const int st = I.period(xhat0 + o * nx, lastu0 + o * nu, ry + o * ny, Zout + o * n, uout + o * nu, &it);
            iters_out[o] = it;
            status_out[o] = st;
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    *seconds = std::chrono::duration<double>(t1 - t0).count();
    int64_t tot = 0;
    for (auto& I : inst) tot += I.s.total_iters;
    *total_admm_iters = tot;
    return 0;
}

// Closed loop WITHOUT any GPU code (bench.py --impl reference records its own trajectory with it): plant = model
// x(k+1) = A x + Bu u, y = C x; SteadyKalmanFilter with gain Khat on the augmented model (correct, moveinput!, predict:
// src/plot_sim.jl:291-311, src/estimator/kalman.jl:284-309); the controller is the per-period path above.
// Plant / observer matrices are ROW-major per instance: A (nxp x nxp), Bu (nxp x nu), C (ny x nxp), Ahat (nx x nx),
// Buhat (nx x nu), Chat (ny x nx), Khat (nx x ny).  ry: (steps x N x ny).  Records xhat0 / lastu0 of every period.
int cpuref_linmpc_closed_loop(int N, int steps, int threads, int nu, int ny, int nx, int nxp, int Hp, int Hc, int neps,
                              const int* nb, const double* E, const double* K, const double* V, const double* B,
                              const double* Ht, const double* Mdiag, const double* U0min, const double* U0max,
                              const double* DUmin, const double* DUmax, const double* Y0min, const double* Y0max,
                              const double* C_umin, const double* C_umax, const double* C_dumin, const double* C_dumax,
                              const double* C_ymin, const double* C_ymax, const double* yop, const double* Ap,
                              const double* Bup, const double* Cp, const double* Ahat, const double* Buhat,
                              const double* Chat, const double* Khat, const double* ry, double* xhat0_rec,
                              double* lastu0_rec, double* u_rec, int32_t* iters_out) {
    Dims d{nu, ny, nx, Hp, Hc, neps, ny * Hp, nu * Hp, nu * Hc, nu * Hc + neps};
    const int n = d.n;
    const std::vector<int> blk = block_of_step(Hp, Hc, nb);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < N; ++i) {
        Inst I;
        bind(I, d, i, E, K, V, B, Mdiag, U0min, U0max, DUmin, DUmax, Y0min, Y0max, yop);
        I.setup(blk, Ht + (size_t)i * n * n, C_umin, C_umax, C_dumin, C_dumax, C_ymin, C_ymax);
        const double* A = Ap + (size_t)i * nxp * nxp;
        const double* Bu = Bup + (size_t)i * nxp * nu;
        const double* C = Cp + (size_t)i * ny * nxp;
        const double* Ah = Ahat + (size_t)i * nx * nx;
        const double* Bh = Buhat + (size_t)i * nx * nu;
        const double* Ch = Chat + (size_t)i * ny * nx;
        const double* Kh = Khat + (size_t)i * nx * ny;
        std::vector<double> x(nxp, 0.0), xn(nxp), xh(nx, 0.0), xhn(nx), y(ny), lu(nu, 0.0), u(nu), v(ny);
        for (int t = 0; t < steps; ++t) {
            const size_t o = (size_t)t * N + i;
            for (int a = 0; a < ny; ++a) {
                double s = 0;
                for (int b2 = 0; b2 < nxp; ++b2) s += C[(size_t)a * nxp + b2] * x[b2];
                y[a] = s;  // (operating points are zero in the synthetic workloads: y = y0)
            }
            // preparestate!: x̂ <- x̂ + K̂ (y0m - Ĉ x̂)
            for (int a = 0; a < ny; ++a) {
                double s = y[a];
                for (int b2 = 0; b2 < nx; ++b2) s -= Ch[(size_t)a * nx + b2] * xh[b2];
                v[a] = s;
            }
            for (int a = 0; a < nx; ++a) {
                double s = xh[a];
                for (int b2 = 0; b2 < ny; ++b2) s += Kh[(size_t)a * ny + b2] * v[b2];
                xhn[a] = s;
            }
            xh = xhn;
            for (int a = 0; a < nx; ++a) xhat0_rec[o * nx + a] = xh[a];
            for (int a = 0; a < nu; ++a) lastu0_rec[o * nu + a] = lu[a];
            int it = 0;
            I.period(xh.data(), lu.data(), ry + o * ny, nullptr, u.data(), &it);
            iters_out[o] = it;
            for (int a = 0; a < nu; ++a) u_rec[o * nu + a] = u[a];
            // plant and observer prediction
            for (int a = 0; a < nxp; ++a) {
                double s = 0;
                for (int b2 = 0; b2 < nxp; ++b2) s += A[(size_t)a * nxp + b2] * x[b2];
                for (int b2 = 0; b2 < nu; ++b2) s += Bu[(size_t)a * nu + b2] * u[b2];
                xn[a] = s;
            }
            x = xn;
            for (int a = 0; a < nx; ++a) {
                double s = 0;
                for (int b2 = 0; b2 < nx; ++b2) s += Ah[(size_t)a * nx + b2] * xh[b2];
                for (int b2 = 0; b2 < nu; ++b2) s += Bh[(size_t)a * nu + b2] * u[b2];
                xhn[a] = s;
            }
            xh = xhn;
            lu = u;
        }
    }
    return 0;
}

// Linear MovingHorizonEstimator, moving-window periods (src/estimator/mhe/execute.jl:44-55, 419-457, 576-617 and
// mhe/transcription.jl:732-781): per period the reference rebuilds H̃ = 2 (ẼZ' M̂ ẼZ + Ñ) with the new arrival
// covariance (M̂ = blockdiag(invP̄, invR̂_He)), the linear term and the right-hand sides, hands them to the solver
// (same sparsity pattern: OSQP update_P -- scaling kept, KKT refactored, warm start kept) and solves.
// Inputs per instance (row-major): EZ (nEZ x n, = [ex̄; Ẽ]), Rdiag (nEZ - nxh: diagonal of invR̂_He), Ntdiag (n: diagonal of Ñ),
// A (m x n: the finite rows of the constraint matrix, constant once the window is full); per period and instance:
// invP (nxh x nxh), FZ (nEZ), b (m), Zs (n: warm start).  Timed: everything from the Hessian rebuild to the solution.
int cpuref_mhe_run(int N, int steps, int threads, int n, int nEZ, int nxh, int m, const double* EZ, const double* Rdiag,
                   const double* Ntdiag, const double* A, const double* invP, const double* FZ, const double* b,
                   const double* Zs, double* Zout, int32_t* iters_out, int32_t* status_out, double* seconds,
                   int64_t* total_admm_iters) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    std::vector<Solver> solvers(N);
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel
    {
        std::vector<double> H((size_t)n * n), MEZ((size_t)nEZ * n), q(n), lvec(m, -INFINITY), uvec(m), zsol(n);
#pragma omp for schedule(dynamic, 1)
        for (int i = 0; i < N; ++i) {
            Solver& s = solvers[i];
            const double* EZi = EZ + (size_t)i * nEZ * n;
            const double* Ai = A + (size_t)i * m * n;
            for (int t = 0; t < steps; ++t) {
                const size_t o = (size_t)t * N + i;
                const double* Pi = invP + o * nxh * nxh;
                const double* Fz = FZ + o * nEZ;
                // M̂ ẼZ  (the first nxh rows through invP̄, the rest through the diagonal invR̂)
                for (int r = 0; r < nxh; ++r)
                    for (int j = 0; j < n; ++j) {
                        double a = 0;
                        for (int k = 0; k < nxh; ++k) a += Pi[(size_t)r * nxh + k] * EZi[(size_t)k * n + j];
                        MEZ[(size_t)r * n + j] = a;
                    }
                for (int r = nxh; r < nEZ; ++r) {
                    const double w = Rdiag[(size_t)i * (nEZ - nxh) + (r - nxh)];
                    for (int j = 0; j < n; ++j) MEZ[(size_t)r * n + j] = w * EZi[(size_t)r * n + j];
                }
                // H̃ = 2 (ẼZ' M̂ ẼZ + Ñ),  q̃ = 2 (M̂ ẼZ)' FZ
                std::fill(H.begin(), H.end(), 0.0);
                for (int r = 0; r < nEZ; ++r) {
                    const double* er = EZi + (size_t)r * n;
                    const double* mr = &MEZ[(size_t)r * n];
                    for (int a = 0; a < n; ++a) {
                        const double ea = er[a];
                        if (ea == 0.0) continue;
                        double* ha = &H[(size_t)a * n];
                        for (int c = 0; c < n; ++c) ha[c] += ea * mr[c];
                    }
                }
                for (int a = 0; a < n; ++a) {
                    for (int c = 0; c < n; ++c) H[(size_t)a * n + c] *= 2.0;
                    H[(size_t)a * n + a] += 2.0 * Ntdiag[(size_t)i * n + a];
                }
                for (int j = 0; j < n; ++j) q[j] = 0.0;
                for (int r = 0; r < nEZ; ++r)
                    for (int j = 0; j < n; ++j) q[j] += MEZ[(size_t)r * n + j] * Fz[r];
                for (int j = 0; j < n; ++j) q[j] *= 2.0;
                if (t == 0) {
                    setup(s, n, m, H.data(), Ai);
                } else {  // update_P: same scaling, new factorisation, iterates kept
                    for (int a = 0; a < n; ++a)
                        for (int c = 0; c < n; ++c) s.P[(size_t)a * n + c] = s.c * s.D[a] * s.D[c] * H[(size_t)a * n + c];
                    factor(s);
                }
                for (int r = 0; r < m; ++r) uvec[r] = b[o * m + r];
                int it = 0;
                const int st = solve(s, q.data(), lvec.data(), uvec.data(), Zs + o * n, zsol.data(), &it);
                for (int j = 0; j < n; ++j) Zout[o * n + j] = zsol[j];
                iters_out[o] = it;
                status_out[o] = st;
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
