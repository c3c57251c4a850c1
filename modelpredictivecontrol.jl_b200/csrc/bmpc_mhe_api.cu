// Host side of the batched linear MovingHorizonEstimator (bmhe_* entry points of include/bmpc.h).
// The handle owns the data windows and the arrival covariance of every instance
// (estim.Y0m, .U0, .D0, .X̂0_old, .x̂0arr_old, .P̂arr_old, cov.invP̄, Nk: reference
// src/estimator/mhe/construct.jl:136-145); the row tables are recompiled whenever the window
// length Nk changes (growing phase), exactly the truncation the reference does in trunc_predmat.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "bmpc_host_util.h"
#include "bmpc_mhe.cuh"

using bmpc_host::DevBuf;
using bmpc_host::even;
using bmpc_host::fail;

struct bmhe_handle {
    bmhe_dims d;
    long NM;
    int nx, nu, nym, nd, He, neps, nZfull;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    bool io_async = false;  // bmhe_set_stream(..., sync = 0): no stream synchronisation after a period (device-resident callers)
    int num_sms = 148;
    bool have_predmat = false, have_cov = false, have_con = false;
    int Nk = 0, compiled_Nk = -1;
    double Cwt = 0.0;
    DevBuf<double> E, EX, G, GX, J, JX, B, BX, A, Cm, Qc, Rm, rinv, Qinv, P0, Rinvd;
    bool R_dense = false;
    DevBuf<double> Y0m, U0, D0, X0old, x0arr, Parr, invP, Z, xhat0, lastu0;
    DevBuf<double> xmin, xmax, wmin, wmax, vmin, vmax;
    std::vector<double> cx_min, cx_max, cw_min, cw_max, cv_min, cv_max;
    std::vector<unsigned char> fin;  // finiteness of x/w/v min/max per component (shared by all instances)
    // compiled per Nk
    bmpc::RowTables rt{};
    DevBuf<int> t_si1, t_si2, t_sch, t_varptr, t_varrow, t_varsgn, t_dbrmax, t_dbrmin, t_drbase, t_drsrc, t_kind,
        t_bidx, t_pdsrc;
    DevBuf<short> t_pi, t_pj;
    DevBuf<double> t_sig, t_c, Pd;
    bmpc::MheLayout L{};
    int nz = 0, n = 0, smem_bytes = 0, grid = 0, nPd = 0;
    // io staging
    DevBuf<double> y0m, d0, u0, Jv, Vhat, X0;
    DevBuf<int> status, iters;
    DevBuf<unsigned int> counter;
    DevBuf<double> Hscratch, lam_ws;
    DevBuf<int> ws_flag, t_wsmap;
    bool warm_start = true;
    bool two_ctas = false;
    int64_t launches = 0;
};

namespace {

__global__ void k_mhe_gather(const double* __restrict__ E, long sE, int ldE, const double* __restrict__ EX, long sEX,
                             int ldEX, double* __restrict__ Pd, long sPd, const int* __restrict__ src, int nDb, int nz,
                             int nXsrc, long tot) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= tot) return;
    const long per = (long)nDb * nz;
    const long inst = e / per;
    const int rem = (int)(e - inst * per);
    const int j = rem / nDb, k = rem - j * nDb;
    const int s = src[k];
    (void)nXsrc;
    Pd[inst * sPd + rem] = s >= 0 ? EX[inst * sEX + s + (long)ldEX * j] : E[inst * sE + (-s - 1) + (long)ldE * j];
}

int compile_rows(bmhe_handle* h, int Nk) {
    const int nx = h->nx, nym = h->nym, neps = h->neps;
    const int nz = nx * (1 + Nk), n = nz + neps;
    const unsigned char* f = h->fin.data();  // [xmin nx | xmax nx | wmin nx | wmax nx | vmin nym | vmax nym]
    auto soft = [&](const std::vector<double>& v, int k) { return neps ? v[k] : 0.0; };
    std::vector<int> s_i1, s_i2, s_ch, kind, bidx, dr_base, dr_src, db_rmax, db_rmin, pd_src;
    std::vector<double> sig, cc;
    auto add_sparse = [&](int var, double sg, double c, int kd, int bi) {
        s_i1.push_back(var); s_i2.push_back(-1); s_ch.push_back(-1);
        sig.push_back(sg); cc.push_back(c); kind.push_back(kd); bidx.push_back(bi);
    };
    for (int k = 0; k < nx; ++k) {
        if (f[k]) add_sparse(k, -1.0, soft(h->cx_min, k), 0, k);
        if (f[nx + k]) add_sparse(k, +1.0, soft(h->cx_max, k), 0, k);
    }
    for (int b = 0; b < Nk; ++b)
        for (int k = 0; k < nx; ++k) {
            if (f[2 * nx + k]) add_sparse(nx + b * nx + k, -1.0, soft(h->cw_min, k), 1, k);
            if (f[3 * nx + k]) add_sparse(nx + b * nx + k, +1.0, soft(h->cw_max, k), 1, k);
        }
    const int nS = (int)s_i1.size();
    auto add_dense = [&](int src, bool hasmin, bool hasmax, double cmin, double cmax, int kd, int bi) {
        if (!hasmin && !hasmax) return;
        const int base = (int)pd_src.size();
        pd_src.push_back(src);
        db_rmax.push_back(-1);
        db_rmin.push_back(-1);
        for (int side = 0; side < 2; ++side) {
            if (!(side ? hasmax : hasmin)) continue;
            (side ? db_rmax : db_rmin)[base] = nS + (int)dr_base.size();
            dr_base.push_back(base);
            dr_src.push_back(0);
            sig.push_back(side ? +1.0 : -1.0);
            cc.push_back(side ? cmax : cmin);
            kind.push_back(kd);
            bidx.push_back(bi);
        }
    };
    for (int t = 0; t < nx * Nk; ++t)
        add_dense(t, f[t % nx], f[nx + t % nx], soft(h->cx_min, t % nx), soft(h->cx_max, t % nx), 2, t);
    for (int t = 0; t < nym * Nk; ++t)
        add_dense(-t - 1, f[4 * nx + t % nym], f[4 * nx + nym + t % nym], soft(h->cv_min, t % nym), soft(h->cv_max, t % nym), 3, t);
    const int nDr = (int)dr_base.size(), nDb = (int)pd_src.size();
    // no eps >= 0 row: the softness weights are non-negative, so eps < 0 is never optimal (see bmpc_set_constraints)
    const int m = nS + nDr;
    // row map of the one-block window shift (moving window, Nk = He on both sides): the arrival rows take over the
    // multipliers of the old X̂ block 0 (same bounds, same state), every Ŵ / X̂ / V̂ block b those of the old block b + 1
    std::vector<int> ws_map(std::max(m, 1), -1);
    {
        int nA = 0, nWb = 0, nVb = 0;
        for (int k = 0; k < nx; ++k) { nA += f[k] + f[nx + k]; nWb += f[2 * nx + k] + f[3 * nx + k]; }
        for (int k = 0; k < nym; ++k) nVb += f[4 * nx + k] + f[4 * nx + nym + k];
        const int nXb = nA, offW = nA, offX = offW + Nk * nWb, offV = offX + Nk * nXb;
        for (int r = 0; r < nA; ++r) ws_map[r] = offX + r;
        for (int b = 0; b + 1 < Nk; ++b) {
            for (int r = 0; r < nWb; ++r) ws_map[offW + b * nWb + r] = offW + (b + 1) * nWb + r;
            for (int r = 0; r < nXb; ++r) ws_map[offX + b * nXb + r] = offX + (b + 1) * nXb + r;
            for (int r = 0; r < nVb; ++r) ws_map[offV + b * nVb + r] = offV + (b + 1) * nVb + r;
        }
        if (offV + Nk * nVb != m) return fail(BMPC_ERR_STATE, "internal: MHE row blocks do not add up (%d != %d)", offV + Nk * nVb, m);
    }
    std::vector<int> var_ptr(nz + 1, 0), var_row(nS), var_sgn(nS, 1);
    for (int g = 0; g < nS; ++g) var_ptr[s_i1[g] + 1]++;
    for (int j = 0; j < nz; ++j) var_ptr[j + 1] += var_ptr[j];
    {
        std::vector<int> cnt(nz, 0);
        for (int g = 0; g < nS; ++g) var_row[var_ptr[s_i1[g]] + cnt[s_i1[g]]++] = g;
    }
    std::vector<short> pi, pj;
    for (int i = 0; i < nz; ++i)
        for (int j = 0; j <= i; ++j) {
            pi.push_back((short)i);
            pj.push_back((short)j);
        }
    cudaStream_t s = h->stream;
    auto nz1 = [](std::vector<int>& v) { if (v.empty()) v.push_back(0); };
    nz1(s_i1); nz1(s_i2); nz1(s_ch); nz1(kind); nz1(bidx); nz1(dr_base); nz1(dr_src); nz1(db_rmax); nz1(db_rmin);
    nz1(pd_src); nz1(var_row); nz1(var_sgn);
    if (sig.empty()) { sig.push_back(0.0); cc.push_back(0.0); }
    CK(h->t_si1.upload(s_i1, s)); CK(h->t_si2.upload(s_i2, s)); CK(h->t_sch.upload(s_ch, s));
    CK(h->t_sig.upload(sig, s)); CK(h->t_c.upload(cc, s)); CK(h->t_varptr.upload(var_ptr, s));
    CK(h->t_varrow.upload(var_row, s)); CK(h->t_varsgn.upload(var_sgn, s)); CK(h->t_dbrmax.upload(db_rmax, s));
    CK(h->t_dbrmin.upload(db_rmin, s)); CK(h->t_drbase.upload(dr_base, s)); CK(h->t_drsrc.upload(dr_src, s));
    CK(h->t_kind.upload(kind, s)); CK(h->t_bidx.upload(bidx, s)); CK(h->t_pdsrc.upload(pd_src, s));
    CK(h->t_pi.upload(pi, s)); CK(h->t_pj.upload(pj, s));
    CK(h->t_wsmap.upload(ws_map, s));
    CK(h->lam_ws.alloc((size_t)h->d.N * even(std::max(m, 1))));
    CK(h->ws_flag.alloc((size_t)h->d.N));
    CK(cudaMemsetAsync(h->ws_flag.p, 0, (size_t)h->d.N * sizeof(int), s));
    if (const char* e = getenv("BMPC_WARM")) h->warm_start = atoi(e) != 0;
    bmpc::RowTables& rt = h->rt;
    rt.nS = nS; rt.nDr = nDr; rt.nDb = nDb; rt.m = m;
    rt.s_i1 = h->t_si1.p; rt.s_i2 = h->t_si2.p; rt.s_ch = h->t_sch.p; rt.row_sig = h->t_sig.p; rt.row_c = h->t_c.p;
    rt.var_ptr = h->t_varptr.p; rt.var_row = h->t_varrow.p; rt.var_sgn = h->t_varsgn.p; rt.db_rmax = h->t_dbrmax.p;
    rt.db_rmin = h->t_dbrmin.p; rt.dr_base = h->t_drbase.p; rt.dr_src = h->t_drsrc.p; rt.pair_i = h->t_pi.p;
    rt.pair_j = h->t_pj.p;
    h->nz = nz;
    h->n = n;
    h->nPd = even(std::max(nDb * nz, 2));
    CK(h->Pd.alloc((size_t)h->NM * h->nPd));
    if (nDb > 0) {
        const long tot = (long)h->NM * nDb * nz;
        k_mhe_gather<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(h->E.p, h->d.shared_model ? 0 : (long)nym * h->He * h->nZfull,
                                                                  nym * h->He, h->EX.p, h->d.shared_model ? 0 : (long)nx * h->He * h->nZfull,
                                                                  nx * h->He, h->Pd.p, (long)h->nPd, h->t_pdsrc.p, nDb, nz, nx * Nk, tot);
        h->launches++;
        CK(cudaGetLastError());
    }
    // shared-memory layout
    bmpc::MheLayout& L = h->L;
    int o = 0;
    auto take = [&](int cnt) { int at = o; o += even(std::max(cnt, 1)); return at; };
    const int nYm = nym * h->He, nXm = nx * h->He, nq = std::max(nx, nym);
    // tuning switch BMHE_TWO_CTAS=1: leave the rebuilt Hessian in an L2-resident scratch slice per CTA and cap the kernel
    // at 128 registers so that two CTAs fit on an SM.  Measured on C3 (8192 x n = 128): 115 k estimates/s against 122 k
    // for one CTA per SM with the Hessian in shared memory -- off by default.
    const int npairs = nz * (nz + 1) / 2;
    h->two_ctas = false;
    if (const char* e = getenv("BMHE_TWO_CTAS")) h->two_ctas = atoi(e) != 0;
    L.Hv = take(h->two_ctas ? 0 : npairs);
    L.Phi = take(std::max(n * (n + 1) / 2, nXm + h->nd * (h->He + 1)));
    L.x = take(n); L.xb = take(n); L.q = take(n); L.rd = take(n); L.rhs = take(n); L.dx = take(n); L.invd = take(n);
    L.yb = take(nDb); L.ybd = take(nDb); L.wd = take(nDb);
    L.s = take(m); L.lam = take(m); L.h = take(m); L.rp = take(m); L.t = take(m); L.ds = take(m); L.dl = take(m);
    L.F = take(nYm); L.FX = take(nXm); L.wrow = take(nYm); L.RF = take(nYm);
    L.P = take(nx * nx); L.P2 = take(2 * nq * nq + nx * nym); L.K = take(nx * nym); L.M = take(nym * nym + nq * nq);
    L.red = take(40);
    L.total = o;
    h->smem_bytes = o * 8;
    int max_optin = 0;
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->d.device));
    if (h->smem_bytes > max_optin)
        return fail(BMPC_ERR_UNSUPPORTED, "MHE window too large for one CTA's shared memory (%d B): n = %d", h->smem_bytes, n);
    int occ = 0;
    if (h->two_ctas) {
        CK(bmpc_host::raise_dyn_smem(reinterpret_cast<const void*>(bmpc::mhe_step_kernel<256, 2>), h->smem_bytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bmpc::mhe_step_kernel<256, 2>, 256, h->smem_bytes));
    } else {
        CK(bmpc_host::raise_dyn_smem(reinterpret_cast<const void*>(bmpc::mhe_step_kernel<256, 1>), h->smem_bytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bmpc::mhe_step_kernel<256, 1>, 256, h->smem_bytes));
    }
    if (occ < 1) return fail(BMPC_ERR_UNSUPPORTED, "MHE kernel does not fit on an SM");
    h->grid = std::max(1, std::min(h->d.N, occ * h->num_sms));
    if (h->two_ctas) CK(h->Hscratch.alloc((size_t)h->grid * even(npairs)));
    h->compiled_Nk = Nk;
    CK(cudaStreamSynchronize(s));
    return BMPC_OK;
}

}  // namespace

extern "C" int bmhe_set_constraints(bmhe_handle* h, const double* xmin, const double* xmax, const double* wmin,
                                    const double* wmax, const double* vmin, const double* vmax, const double* c_x,
                                    const double* c_w, const double* c_v);

// add_data_windows! + (correct_cov!) + initpred! + linconstraint! + optim_objective! + getstate! for the whole batch:
// the body shared by correct_estimate! (direct = true, window input = the stored u0(k-1)) and update_estimate!
// (direct = false, window input = u0(k), already uploaded to h->u0): src/estimator/mhe/execute.jl:44-84.
static int solve_window(bmhe_handle* h, const double* y0m, const double* d0, const double* u_window_dev, double* xhat0,
                        double* Ztilde, double* J, int32_t* status, int32_t* iters, double* Vhat, double* X0) {
    if (!h || !y0m || !xhat0 || !status || !iters) return fail(BMPC_ERR_ARG, "null argument");
    if (!h->have_predmat || !h->have_cov) return fail(BMPC_ERR_STATE, "bmhe_set_predmat and bmhe_set_cov must be called first");
    if (h->nd > 0 && !d0) return fail(BMPC_ERR_ARG, "d0 is required when nd > 0");
    CK(cudaSetDevice(h->d.device));
    if (!h->have_con) {
        int rc = bmhe_set_constraints(h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc != BMPC_OK) return rc;
    }
    const int moving = h->Nk + 1 > h->He;
    const int Nk = std::min(h->Nk + 1, h->He);
    if (Nk != h->compiled_Nk) {
        int rc = compile_rows(h, Nk);
        if (rc != BMPC_OK) return rc;
    }
    cudaStream_t s = h->stream;
    const size_t N = h->d.N, nx = h->nx, nym = h->nym, nd = h->nd, He = h->He;
    CK(cudaMemcpyAsync(h->y0m.p, y0m, N * nym * 8, cudaMemcpyDefault, s));
    if (nd) CK(cudaMemcpyAsync(h->d0.p, d0, N * nd * 8, cudaMemcpyDefault, s));
    bmpc::StepParams P{};
    P.N = h->d.N; P.nz = h->nz; P.n = h->n; P.neps = h->neps; P.max_iter = h->d.max_iter; P.tol = h->d.tol;
    P.tol_mu = h->d.tol * 1e-3; P.rt = h->rt;
    bmpc::MheParams Q{};
    Q.N = h->d.N; Q.nx = h->nx; Q.nu = h->nu; Q.nym = h->nym; Q.nd = h->nd; Q.He = h->He; Q.Nk = Nk; Q.neps = h->neps;
    Q.moving = moving; Q.direct = h->d.direct ? 1 : 0; Q.ldE = h->nym * h->He; Q.ldEX = h->nx * h->He;
    const long sh = h->d.shared_model ? 0 : 1;
    const long nZ = h->nZfull;
    Q.sE = sh * nym * He * nZ; Q.sEX = sh * nx * He * nZ; Q.sG = sh * nym * He * h->nu * He; Q.sGX = sh * nx * He * h->nu * He;
    Q.sJ = sh * nym * He * nd * (He + 1); Q.sJX = sh * nx * He * nd * (He + 1); Q.sB = sh * nym * He; Q.sBX = sh * nx * He;
    Q.sCm = sh * nym * nx; Q.sCov = sh;
    Q.E = h->E.p; Q.EX = h->EX.p; Q.G = h->G.p; Q.GX = h->GX.p; Q.J = h->J.p; Q.JX = h->JX.p; Q.B = h->B.p; Q.BX = h->BX.p;
    Q.Cm = h->Cm.p; Q.Rm = h->Rm.p; Q.rinv = h->rinv.p; Q.Qinv = h->Qinv.p; Q.Cwt = h->Cwt;
    Q.Rinvd = h->R_dense ? h->Rinvd.p : nullptr;
    Q.Y0m = h->Y0m.p; Q.U0 = h->U0.p; Q.D0 = h->D0.p; Q.X0old = h->X0old.p; Q.x0arr = h->x0arr.p; Q.Parr = h->Parr.p;
    Q.invP = h->invP.p; Q.Z = h->Z.p; Q.xhat0 = h->xhat0.p; Q.lastu0 = const_cast<double*>(u_window_dev);
    Q.xmin = h->xmin.p; Q.xmax = h->xmax.p; Q.wmin = h->wmin.p; Q.wmax = h->wmax.p; Q.vmin = h->vmin.p; Q.vmax = h->vmax.p;
    Q.row_kind = h->t_kind.p; Q.row_bidx = h->t_bidx.p; Q.Pd = h->Pd.p; Q.sPd = sh * h->nPd;
    Q.y0m = h->y0m.p; Q.d0 = h->d0.p; Q.J_out = h->Jv.p; Q.Vhat_out = Vhat ? h->Vhat.p : nullptr; Q.X0_out = X0 ? h->X0.p : nullptr;
    Q.status = h->status.p; Q.iters = h->iters.p; Q.L = h->L;
    if (!h->counter.p) CK(h->counter.alloc(1));
    CK(cudaMemsetAsync(h->counter.p, 0, sizeof(unsigned int), s));
    Q.counter = h->counter.p;
    Q.lam_ws = h->lam_ws.p; Q.ws_flag = h->ws_flag.p; Q.ws_map = h->t_wsmap.p; Q.ws_stride = even(std::max(h->rt.m, 1));
    Q.use_ws = h->warm_start ? 1 : 0;
    Q.Hscratch = h->two_ctas ? h->Hscratch.p : nullptr;
    Q.sHs = even(h->nz * (h->nz + 1) / 2);
    if (h->two_ctas)
        bmpc::mhe_step_kernel<256, 2><<<h->grid, 256, h->smem_bytes, s>>>(P, Q);
    else
        bmpc::mhe_step_kernel<256, 1><<<h->grid, 256, h->smem_bytes, s>>>(P, Q);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(BMPC_ERR_CUDA, "MHE kernel launch failed: %s", cudaGetErrorString(le));
    h->launches++;
    h->Nk = Nk;
    CK(cudaMemcpyAsync(xhat0, h->xhat0.p, N * nx * 8, cudaMemcpyDefault, s));
    if (Ztilde) CK(cudaMemcpyAsync(Ztilde, h->Z.p, N * (h->neps + nx * (1 + He)) * 8, cudaMemcpyDefault, s));
    if (J) CK(cudaMemcpyAsync(J, h->Jv.p, N * 8, cudaMemcpyDefault, s));
    if (Vhat) CK(cudaMemcpyAsync(Vhat, h->Vhat.p, N * nym * He * 8, cudaMemcpyDefault, s));
    if (X0) CK(cudaMemcpyAsync(X0, h->X0.p, N * nx * He * 8, cudaMemcpyDefault, s));
    CK(cudaMemcpyAsync(status, h->status.p, N * 4, cudaMemcpyDefault, s));
    CK(cudaMemcpyAsync(iters, h->iters.p, N * 4, cudaMemcpyDefault, s));
    if (!h->io_async) CK(cudaStreamSynchronize(s));
    return BMPC_OK;
}


extern "C" {

int bmhe_create(bmhe_handle** out, const bmhe_dims* dims) {
    if (!out || !dims) return fail(BMPC_ERR_ARG, "null argument");
    const bmhe_dims& d = *dims;
    if (d.N < 1 || d.nu < 1 || d.nym < 1 || d.nd < 0 || d.nxhat < 1 || d.He < 1 || (d.neps != 0 && d.neps != 1))
        return fail(BMPC_ERR_ARG, "invalid dimensions");
    if (d.nxhat > 32 || d.nym > 16) return fail(BMPC_ERR_UNSUPPORTED, "nxhat <= 32 and nym <= 16 are supported");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(BMPC_ERR_CUDA, "no CUDA device: libbmpc has no CPU fallback (%s)", cudaGetErrorString(e));
    if (d.device < 0 || d.device >= ndev) return fail(BMPC_ERR_ARG, "device %d out of range", d.device);
    CK(cudaSetDevice(d.device));
    bmhe_handle* h = new bmhe_handle();
    h->d = d;
    if (h->d.max_iter <= 0) h->d.max_iter = 50;
    if (!(h->d.tol > 0)) h->d.tol = 1e-11;
    h->nx = d.nxhat; h->nu = d.nu; h->nym = d.nym; h->nd = d.nd; h->He = d.He; h->neps = d.neps;
    h->nZfull = d.nxhat * (1 + d.He);
    h->NM = d.shared_model ? 1 : d.N;
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, d.device);
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return fail(BMPC_ERR_CUDA, "cudaStreamCreate failed");
    }
    h->stream = h->own_stream;
    const size_t N = d.N, nx = d.nxhat, He = d.He;
    cudaError_t a = cudaSuccess;
    auto A = [&](cudaError_t r) { if (a == cudaSuccess) a = r; };
    A(h->Y0m.alloc(N * d.nym * He)); A(h->U0.alloc(N * d.nu * He)); A(h->D0.alloc(N * std::max(d.nd, 1) * (He + 1)));
    A(h->X0old.alloc(N * nx * He)); A(h->x0arr.alloc(N * nx)); A(h->Parr.alloc(N * nx * nx)); A(h->invP.alloc(N * nx * nx));
    A(h->Z.alloc(N * (d.neps + nx * (1 + He)))); A(h->xhat0.alloc(N * nx)); A(h->lastu0.alloc(N * d.nu));
    A(h->y0m.alloc(N * d.nym)); A(h->d0.alloc(N * std::max(d.nd, 1))); A(h->u0.alloc(N * d.nu)); A(h->Jv.alloc(N));
    A(h->Vhat.alloc(N * d.nym * He)); A(h->X0.alloc(N * nx * He)); A(h->status.alloc(N)); A(h->iters.alloc(N));
    if (a != cudaSuccess) {
        bmhe_destroy(h);
        return fail(BMPC_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(a));
    }
    const int nfin = 4 * d.nxhat + 2 * d.nym;
    h->fin.assign(nfin, 0);
    h->cx_min.assign(nx, 0.0); h->cx_max.assign(nx, 0.0); h->cw_min.assign(nx, 0.0); h->cw_max.assign(nx, 0.0);
    h->cv_min.assign(d.nym, 0.0); h->cv_max.assign(d.nym, 0.0);
    *out = h;
    return BMPC_OK;
}

int bmhe_destroy(bmhe_handle* h) {
    if (!h) return BMPC_OK;
    cudaSetDevice(h->d.device);
    cudaDeviceSynchronize();
    DevBuf<double>* bufs[] = {&h->E, &h->EX, &h->G, &h->GX, &h->J, &h->JX, &h->B, &h->BX, &h->A, &h->Cm, &h->Qc, &h->Rm,
                              &h->rinv, &h->Qinv, &h->P0, &h->Y0m, &h->U0, &h->D0, &h->X0old, &h->x0arr, &h->Parr,
                              &h->invP, &h->Z, &h->xhat0, &h->lastu0, &h->xmin, &h->xmax, &h->wmin, &h->wmax, &h->vmin,
                              &h->vmax, &h->t_sig, &h->t_c, &h->Pd, &h->y0m, &h->d0, &h->u0, &h->Jv, &h->Vhat, &h->X0};
    for (auto* b : bufs) b->release();
    DevBuf<int>* ib[] = {&h->t_si1, &h->t_si2, &h->t_sch, &h->t_varptr, &h->t_varrow, &h->t_varsgn, &h->t_dbrmax,
                         &h->t_dbrmin, &h->t_drbase, &h->t_drsrc, &h->t_kind, &h->t_bidx, &h->t_pdsrc, &h->status, &h->iters};
    for (auto* b : ib) b->release();
    h->t_pi.release();
    h->t_pj.release();
    h->counter.release();
    h->Hscratch.release();
    h->lam_ws.release();
    h->ws_flag.release();
    h->t_wsmap.release();
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return BMPC_OK;
}

int bmhe_set_predmat(bmhe_handle* h, const double* E, const double* G, const double* J, const double* B,
                     const double* EX, const double* GX, const double* JX, const double* BX) {
    if (!h || !E || !G || !B || !EX || !GX || !BX) return fail(BMPC_ERR_ARG, "null argument");
    if (h->nd > 0 && (!J || !JX)) return fail(BMPC_ERR_ARG, "J and JX are required when nd > 0");
    CK(cudaSetDevice(h->d.device));
    const size_t NM = h->NM, nx = h->nx, nu = h->nu, nym = h->nym, nd = h->nd, He = h->He, nZ = h->nZfull;
    cudaStream_t s = h->stream;
    CK(h->E.upload(E, NM * nym * He * nZ, s)); CK(h->EX.upload(EX, NM * nx * He * nZ, s));
    CK(h->G.upload(G, NM * nym * He * nu * He, s)); CK(h->GX.upload(GX, NM * nx * He * nu * He, s));
    CK(h->B.upload(B, NM * nym * He, s)); CK(h->BX.upload(BX, NM * nx * He, s));
    if (nd) {
        CK(h->J.upload(J, NM * nym * He * nd * (He + 1), s));
        CK(h->JX.upload(JX, NM * nx * He * nd * (He + 1), s));
    }
    CK(cudaStreamSynchronize(s));
    h->have_predmat = true;
    h->compiled_Nk = -1;
    return BMPC_OK;
}

int bmhe_set_cov(bmhe_handle* h, const double* Ahat, const double* Cmhat, const double* P0, const double* Qhat,
                 const double* Rhat, double Cwt) {
    if (!h || !Ahat || !Cmhat || !P0 || !Qhat || !Rhat) return fail(BMPC_ERR_ARG, "null argument");
    if (h->neps && !(std::isfinite(Cwt) && Cwt >= 0)) return fail(BMPC_ERR_ARG, "Cwt must be finite and >= 0 when neps = 1");
    CK(cudaSetDevice(h->d.device));
    const size_t NM = h->NM, nx = h->nx, nym = h->nym;
    cudaStream_t s = h->stream;
    // inverse covariances on the host (one-off).  A diagonal Rhat (the reference default, kalman.jl:166-171) keeps the
    // per-row weights; a non-diagonal one switches the kernel to the dense R̂^-1 path
    std::vector<double> rinv(NM * nym), Qinv(NM * nx * nx), Rinvd(NM * nym * nym);
    // inverse of a small SPD matrix by Gauss-Jordan (column-major n x n); false if a pivot is not positive
    auto spd_inverse = [](const double* Msrc, size_t n, double* out) {
        std::vector<double> M(Msrc, Msrc + n * n), I(n * n, 0.0);
        for (size_t k = 0; k < n; ++k) I[k + n * k] = 1.0;
        for (size_t k = 0; k < n; ++k) {
            const double piv = M[k + n * k];
            if (!(piv > 0)) return false;
            for (size_t c = 0; c < n; ++c) { M[k + n * c] /= piv; I[k + n * c] /= piv; }
            for (size_t r = 0; r < n; ++r) {
                if (r == k) continue;
                const double fct = M[r + n * k];
                for (size_t c = 0; c < n; ++c) { M[r + n * c] -= fct * M[k + n * c]; I[r + n * c] -= fct * I[k + n * c]; }
            }
        }
        std::copy(I.begin(), I.end(), out);
        return true;
    };
    bool r_dense = false;
    for (size_t i = 0; i < NM; ++i) {
        const double* R = Rhat + i * nym * nym;
        for (size_t a = 0; a < nym; ++a)
            for (size_t b = 0; b < nym; ++b) {
                if (a != b && R[a + nym * b] != 0.0) r_dense = true;
                if (a == b) {
                    if (!(R[a + nym * a] > 0)) return fail(BMPC_ERR_ARG, "Rhat is not positive definite");
                    rinv[i * nym + a] = 1.0 / R[a + nym * a];
                }
            }
        if (!spd_inverse(R, nym, Rinvd.data() + i * nym * nym)) return fail(BMPC_ERR_ARG, "Rhat is not positive definite");
        if (!spd_inverse(Qhat + i * nx * nx, nx, Qinv.data() + i * nx * nx)) return fail(BMPC_ERR_ARG, "Qhat is not positive definite");
    }
    h->R_dense = r_dense;
    CK(h->Rinvd.upload(Rinvd, s));
    CK(h->A.upload(Ahat, NM * nx * nx, s)); CK(h->Cm.upload(Cmhat, NM * nym * nx, s));
    CK(h->P0.upload(P0, NM * nx * nx, s)); CK(h->Qc.upload(Qhat, NM * nx * nx, s)); CK(h->Rm.upload(Rhat, NM * nym * nym, s));
    CK(h->rinv.upload(rinv, s)); CK(h->Qinv.upload(Qinv, s));
    CK(cudaStreamSynchronize(s));
    h->Cwt = Cwt;
    h->have_cov = true;
    return bmhe_reset(h);
}

int bmhe_set_constraints(bmhe_handle* h, const double* xmin, const double* xmax, const double* wmin, const double* wmax,
                         const double* vmin, const double* vmax, const double* c_x, const double* c_w, const double* c_v) {
    if (!h) return fail(BMPC_ERR_ARG, "null handle");
    CK(cudaSetDevice(h->d.device));
    const int N = h->d.N, nx = h->nx, nym = h->nym;
    const double* arrs[6] = {xmin, xmax, wmin, wmax, vmin, vmax};
    const int lens[6] = {nx, nx, nx, nx, nym, nym};
    std::vector<unsigned char> fin(4 * nx + 2 * nym, 0);
    std::vector<std::vector<double>> vals(6);
    int o = 0;
    for (int a = 0; a < 6; ++a) {
        vals[a].assign((size_t)N * lens[a], (a & 1) ? INFINITY : -INFINITY);
        for (int k = 0; k < lens[a]; ++k, ++o)
            for (int i = 0; i < N; ++i) {
                const double v = arrs[a] ? arrs[a][(size_t)i * lens[a] + k] : ((a & 1) ? INFINITY : -INFINITY);
                if (std::isnan(v)) return fail(BMPC_ERR_ARG, "NaN bound");
                const unsigned char fn = std::isfinite(v) ? 1 : 0;
                if (i == 0) fin[o] = fn;
                else if (fin[o] != fn) return fail(BMPC_ERR_ARG, "all instances must share the +-Inf pattern of their bounds");
                vals[a][(size_t)i * lens[a] + k] = fn ? v : 0.0;
            }
    }
    if (h->Nk > 0 && fin != h->fin) return fail(BMPC_ERR_STATE, "Cannot modify +-Inf constraints after the first step");
    for (const double* cs : {c_x, c_w, c_v})
        if (cs && h->neps)
            for (int k = 0; k < 2 * (cs == c_v ? nym : nx); ++k)
                if (cs[k] < 0) return fail(BMPC_ERR_ARG, "softness weights should be non-negative (mhe/construct.jl setconstraint!)");
    auto setc = [&](std::vector<double>& dst, const double* src, int len) {
        for (int k = 0; k < len; ++k) dst[k] = (src && h->neps) ? src[k] : 0.0;
    };
    setc(h->cx_min, c_x, nx); setc(h->cx_max, c_x ? c_x + nx : nullptr, nx);
    setc(h->cw_min, c_w, nx); setc(h->cw_max, c_w ? c_w + nx : nullptr, nx);
    setc(h->cv_min, c_v, nym); setc(h->cv_max, c_v ? c_v + nym : nullptr, nym);
    cudaStream_t s = h->stream;
    CK(h->xmin.upload(vals[0], s)); CK(h->xmax.upload(vals[1], s)); CK(h->wmin.upload(vals[2], s));
    CK(h->wmax.upload(vals[3], s)); CK(h->vmin.upload(vals[4], s)); CK(h->vmax.upload(vals[5], s));
    CK(cudaStreamSynchronize(s));
    h->fin = fin;
    h->have_con = true;
    h->compiled_Nk = -1;
    return BMPC_OK;
}

int bmhe_reset(bmhe_handle* h) {
    if (!h || !h->have_cov) return fail(BMPC_ERR_STATE, "bmhe_set_cov must be called first");
    CK(cudaSetDevice(h->d.device));
    const size_t N = h->d.N, nx = h->nx, He = h->He;
    cudaStream_t s = h->stream;
    // init_estimate_cov! (execute.jl:2-37): windows NaN, u0(-1) = 0, d0(-1) = 0, P̄ = P̂_0, x̂0 = 0, Nk = 0
    std::vector<double> nanv(N * std::max({(size_t)h->nym * He, (size_t)h->nu * He, nx * He, (size_t)std::max(h->nd, 1) * (He + 1)}), NAN);
    CK(cudaMemcpyAsync(h->Y0m.p, nanv.data(), N * h->nym * He * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->U0.p, nanv.data(), N * h->nu * He * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->X0old.p, nanv.data(), N * nx * He * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(h->D0.p, 0, N * std::max(h->nd, 1) * (He + 1) * 8, s));
    CK(cudaMemsetAsync(h->Z.p, 0, N * (h->neps + nx * (1 + He)) * 8, s));
    CK(cudaMemsetAsync(h->xhat0.p, 0, N * nx * 8, s));
    CK(cudaMemsetAsync(h->x0arr.p, 0, N * nx * 8, s));
    CK(cudaMemsetAsync(h->lastu0.p, 0, N * h->nu * 8, s));
    // P̄ and its inverse (host, one-off)
    std::vector<double> P0(h->NM * nx * nx);
    CK(cudaMemcpyAsync(P0.data(), h->P0.p, P0.size() * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<double> Pall(N * nx * nx), Iall(N * nx * nx);
    for (size_t i = 0; i < N; ++i) {
        const double* P = P0.data() + (h->d.shared_model ? 0 : i) * nx * nx;
        std::vector<double> M(P, P + nx * nx), I(nx * nx, 0.0);
        for (size_t k = 0; k < nx; ++k) I[k + nx * k] = 1.0;
        for (size_t k = 0; k < nx; ++k) {
            const double piv = M[k + nx * k];
            if (!(piv > 0)) return fail(BMPC_ERR_ARG, "P0 is not positive definite");
            for (size_t c = 0; c < nx; ++c) { M[k + nx * c] /= piv; I[k + nx * c] /= piv; }
            for (size_t r = 0; r < nx; ++r) {
                if (r == k) continue;
                const double fct = M[r + nx * k];
                for (size_t c = 0; c < nx; ++c) { M[r + nx * c] -= fct * M[k + nx * c]; I[r + nx * c] -= fct * I[k + nx * c]; }
            }
        }
        std::copy(P, P + nx * nx, Pall.begin() + i * nx * nx);
        std::copy(I.begin(), I.end(), Iall.begin() + i * nx * nx);
    }
    CK(cudaMemcpyAsync(h->Parr.p, Pall.data(), Pall.size() * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->invP.p, Iall.data(), Iall.size() * 8, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    h->Nk = 0;
    return BMPC_OK;
}

int bmhe_correct(bmhe_handle* h, const double* y0m, const double* d0, double* xhat0, double* Ztilde, double* J,
                 int32_t* status, int32_t* iters, double* Vhat, double* X0) {
    if (!h || !xhat0) return fail(BMPC_ERR_ARG, "null argument");
    if (!h->d.direct) {
        // prediction form: preparestate! leaves the estimate untouched (correct_estimate! is empty, execute.jl:44-55)
        CK(cudaSetDevice(h->d.device));
        CK(cudaMemcpyAsync(xhat0, h->xhat0.p, (size_t)h->d.N * h->nx * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return BMPC_OK;
    }
    return solve_window(h, y0m, d0, h->lastu0.p, xhat0, Ztilde, J, status, iters, Vhat, X0);
}

static int launch_update(bmhe_handle* h) {
    cudaStream_t s = h->stream;
    const size_t N = h->d.N, nx = h->nx, nym = h->nym, nq = std::max(nx, nym);
    const long sh = h->d.shared_model ? 0 : 1;
    const int smem = (int)((2 * nx * nx + 2 * nq * nq + 2 * nx * nym + nym * nym + nq * nq) * 8);
    CK(bmpc_host::raise_dyn_smem(reinterpret_cast<const void*>(bmpc::k_mhe_update), smem));
    bmpc::k_mhe_update<<<(unsigned)N, 64, smem, s>>>((int)N, (int)nx, h->nu, (int)nym, h->Nk == h->He ? 1 : 0, h->d.direct ? 1 : 0,
                                                    h->A.p, sh * (long)(nx * nx), h->Qc.p, sh * (long)(nx * nx), h->Cm.p,
                                                    sh * (long)(nym * nx), h->Rm.p, sh * (long)(nym * nym), h->Parr.p, h->invP.p,
                                                    h->lastu0.p, h->u0.p);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(BMPC_ERR_CUDA, "MHE update kernel launch failed: %s", cudaGetErrorString(le));
    h->launches++;
    if (!h->io_async) CK(cudaStreamSynchronize(s));
    return BMPC_OK;
}

int bmhe_update(bmhe_handle* h, const double* u0) {
    if (!h || !u0) return fail(BMPC_ERR_ARG, "null argument");
    if (!h->d.direct) return fail(BMPC_ERR_STATE, "direct = false: updatestate! needs ym and d, call bmhe_update_solve");
    if (h->Nk < 1) return fail(BMPC_ERR_STATE, "bmhe_correct (preparestate!) must be called before bmhe_update");
    CK(cudaSetDevice(h->d.device));
    CK(cudaMemcpyAsync(h->u0.p, u0, (size_t)h->d.N * h->nu * 8, cudaMemcpyDefault, h->stream));
    return launch_update(h);
}

int bmhe_update_solve(bmhe_handle* h, const double* u0, const double* y0m, const double* d0, double* xhat0,
                      double* Ztilde, double* J, int32_t* status, int32_t* iters, double* Vhat, double* X0) {
    if (!h || !u0) return fail(BMPC_ERR_ARG, "null argument");
    if (h->d.direct) return fail(BMPC_ERR_STATE, "direct = true: the window is solved in bmhe_correct; call bmhe_update");
    CK(cudaSetDevice(h->d.device));
    CK(cudaMemcpyAsync(h->u0.p, u0, (size_t)h->d.N * h->nu * 8, cudaMemcpyDefault, h->stream));
    int rc = solve_window(h, y0m, d0, h->u0.p, xhat0, Ztilde, J, status, iters, Vhat, X0);
    if (rc != BMPC_OK) return rc;
    return launch_update(h);
}

int bmhe_set_stream(bmhe_handle* h, void* stream, int32_t sync) {
    if (!h) return fail(BMPC_ERR_ARG, "null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    h->io_async = sync == 0;
    return BMPC_OK;
}

int64_t bmhe_launch_count(bmhe_handle* h) { return h ? h->launches : 0; }

}  // extern "C"
