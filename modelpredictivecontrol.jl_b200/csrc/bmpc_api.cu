// Host side of libbmpc.so: the C ABI declared in include/bmpc.h.
//  - owns all device memory of a batch of LinMPC controllers,
//  - "compiles" the constraint structure (setconstraint!, construct.jl:324-559 and
//    init_matconstraint_mpc, transcription.jl:667-703) into merged row tables in input-level
//    coordinates,
//  - launches the step kernel (bmpc_device.cuh).
// No CPU fallback exists: without a CUDA device bmpc_create fails with BMPC_ERR_CUDA.
#include "bmpc_host_util.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "bmpc_device.cuh"
#include "bmpc_kf.cuh"
#include "bmpc_setup.cuh"
#include "bmpc_model.cuh"
#include "bmpc_warp_registry.h"

namespace bmpc_host {
std::string& last_error() {
    thread_local std::string err;
    return err;
}
}  // namespace bmpc_host

using bmpc_host::DevBuf;
using bmpc_host::even;
using bmpc_host::fail;

struct bmpc_handle {
    bmpc_dims d;
    std::vector<int> nb, blk_of_t, blk_start;
    int nz, n, nY, nU, nHp2, nEv2;
    const double *last_Z = nullptr, *last_xhat0 = nullptr, *last_d0 = nullptr, *last_Dhat0 = nullptr;  // device pointers used by the last step (getinfo)
    long NM;  // number of model copies: N or 1 (shared_model)
    cudaStream_t stream = nullptr, own_stream = nullptr;
    int num_sms = 148;
    bool have_predmat = false, have_weights = false, have_constraints = false, stepped = false;
    bool has_terminal_mats = false, M_dense = false, has_L = false, L_dense = false, route_a = false;
    // model-dependent constants
    DevBuf<double> E, ex, Ht;                         // reference coordinates (kept for getinfo)
    DevBuf<double> Ev, exv, Hv, Lv, Hee, K, V, B, G, J, kx, vx, bx, gx, jx, Mw, Lw, uop, yop;
    DevBuf<int> lv_ok;
    // constraint tables
    bmpc::RowTables rt{};
    DevBuf<int> t_si1, t_si2, t_sch, t_varptr, t_varrow, t_varsgn, t_dbrmax, t_dbrmin, t_drbase, t_drsrc, t_blk,
        t_blkstart, t_pdsrc;
    DevBuf<short> t_pi, t_pj;
    DevBuf<double> t_sig, t_c, sbase, dbound, Pd;
    // augmented model kept from route A (bmpc_set_model) for the MultipleShooting state block (bmpc_get_states)
    DevBuf<double> mA, mBu, mBd, mf;
    bool have_model_mats = false;
    // custom linear constraints (bmpc_set_custom / bmpc_set_custom_bounds)
    int nw = 0, nFw = 0;
    DevBuf<double> Wc, Ewv;
    long sWc = 0;
    std::vector<double> hWmin, hWmax, hCwmin, hCwmax;  // N x nFw bounds (absolute), nFw softness (shared)
    // host copies of the row tables (the warp kernel's position tables are derived from them)
    std::vector<int> hs_i1, hs_i2, hs_ch, hdr_src, hdr_base, hdb_rmax, hdb_rmin;
    std::vector<double> hsig, hcc;
    std::vector<unsigned char> pattern;  // finiteness pattern of the bounds (frozen after first step)
    int nPd2 = 0;
    bool pd_is_ev = false, pd_in_smem = false, has_terminal_rows = false, hv_in_smem = true;
    // io staging
    DevBuf<double> xhat0, lastu0, ry, Rhat_y, Rhat_u, d0, Dhat0, Z, u, Jv, F, qt, r, lastu_prev, Ys, kkt;
    DevBuf<int> status, iters;
    DevBuf<unsigned int> counters;
    // launch geometry
    int team = 0, teams_per_cta = 0, grid = 0, smem_bytes = 0;
    bool dirty = true;      // launch geometry / derived arrays must be rebuilt before the next step
    DevBuf<double> lam_ws;
    DevBuf<int> ws_flag;
    // warp-per-controller kernel (bmpc_warp.cuh)
    const bmpc::WarpEntry* warp = nullptr;
    bmpc::WarpParams wp{};
    DevBuf<double> Gw, HLw;
    DevBuf<int4> w_pinfo;   // per position: row, flags, bound source / shift channel, unit variable
    DevBuf<int> w_upos, w_posrow;
    DevBuf<int> order[2];
    DevBuf<unsigned int> ocnt;
#ifdef BMPC_PHASE_CLK
    DevBuf<long long> clk;  // study builds: per-phase cycle accumulators of step_warp
#endif
    // fused observer (bmpc_set_estimator): model matrices, state x̂0 (handle-owned), corrected estimate of the last step
    bool have_estimator = false;
    int nym = 0;
    DevBuf<double> eA, eBu, eBd, eCm, eDdm, eK, efx, xstate, xcorr, y0m;
    DevBuf<double> kfP, kfQ, kfR;  // time-varying KalmanFilter: P̂, Q̂, R̂ per model (bmpc_set_estimator_cov)
    bool kf_on = false;
    bool e_has_fx = false;
    double* zg[8] = {nullptr};  // peer-mapped gather buffers (bmpc_set_gather / bmpc_set_gather_flags)
    int zg_world = 0, zg_rank = 0;
    unsigned long long* zg_flag[8] = {nullptr};  // pull protocol: every peer's flag array [2 world] (data epochs, ack epochs)
    long zg_row_offset = 0, zg_rows_total = 0;
    int zg_row_off[9] = {0};
    int zg_slots = 1;
    bool zg_pull = false, zg_consumer = false;
    cudaEvent_t zg_ev = nullptr;
    int64_t zg_epoch = 0;   // periods published so far
    DevBuf<int> zg_timeout;
    DevBuf<unsigned int> zg_done;
    int order_cur = 0;       // order[order_cur] drives the next launch
    bool order_valid = false;
    int warm_start = 1;
    bmpc::SmemLayout sm{};
    int64_t launches = 0;
};

namespace {

int choose_team(const bmpc_handle* h) {
    if (h->d.team) return h->d.team;
    if (const char* e = getenv("BMPC_TEAM")) return atoi(e);  // tuning override
    const int n = h->n;
    if (n <= 16) return 16;
    // CTA teams: Hessian build on the FP64 tensor pipe (DMMA), several CTAs per SM.  Throughput follows the number of
    // resident instances per SM, which shared memory caps: two warps per instance with the packed Hessian left in L2
    // (8 CTAs/SM at n = 41) measured 776 k steps/s on C2 against 700 k for four warps with the Hessian staged (6 CTAs/SM)
    if (n <= 48) return 64;
    if (n <= 96) return 128;
    return 256;
}

// shared-memory slice of one team, in doubles (every offset even => 16-byte aligned)
void layout_smem(bmpc_handle* h, bool pd_in_smem, bool hv_in_smem = true) {
    const int n = h->n, nz = h->nz, m = h->rt.m, nDb = h->rt.nDb, nY = h->nY, nx = h->d.nxhat;
    bmpc::SmemLayout& L = h->sm;
    int o = 0;
    auto take = [&](int cnt) {
        int at = o;
        o += even(std::max(cnt, 1));
        return at;
    };
    L.Pd = take(pd_in_smem ? h->nPd2 : 0);
    L.Hv = take(hv_in_smem ? h->nHp2 : 0);
    L.Phi = take(std::max(even(n * (n + 1) / 2), h->nHp2));
    L.x = take(n);
    L.xb = take(n);
    L.q = take(n);
    L.rd = take(n);
    L.rhs = take(n);
    L.dx = take(n);
    L.invd = take(n);
    L.fx = take(nx);
    L.yb = take(nDb);
    L.ybd = take(nDb);
    L.wd = take(nDb);
    L.s = take(m);
    L.lam = take(m);
    L.h = take(m);
    L.rp = take(m);
    L.t = take(std::max(m, nY));
    L.ds = take(std::max(m, nY));
    L.dl = take(std::max(m, nY));
    // F and tY live only until the bound vector h and the gradient q are built (stage 1); t and ds only inside the
    // interior-point solve: they share storage (1.9 KB per instance at C2 = one more resident CTA per SM)
    L.F = L.ds;
    L.tY = L.t;
    L.xhat = take(nx);
    L.lastu = take(h->d.nu);
    L.dd = take(h->d.nd);
    L.Dh = take(h->d.nd * h->d.Hp);
    L.red = take(40);
    L.bar = take(2);
    L.ev = take(h->d.ny);
    L.Fw = take(h->nFw);
    L.cu = take(h->L_dense ? h->nU : 0);
    L.tU = take(h->L_dense ? h->nU : 0);
    L.total = o;
    (void)nz;
}

template <int TEAM>
int configure(bmpc_handle* h) {
    constexpr int CTA = bmpc::CtaThreads<TEAM>::value;
    const size_t pd_bytes = (size_t)h->nPd2 * 8;
    int max_optin = 0;
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->d.device));
    h->team = TEAM;
    h->teams_per_cta = CTA / TEAM;
    // staging policy: Pd in shared memory only when small (it is re-read from L1/L2 otherwise); for CTA teams keep
    // the per-CTA footprint low enough for several CTAs per SM; the packed Hessian leaves shared memory last
    size_t pd_limit = TEAM >= 64 ? 24 * 1024 : 96 * 1024;
    if (const char* e = getenv("BMPC_PD_LIMIT")) pd_limit = (size_t)atol(e);  // tuning override (bytes)
    bool in_smem = h->rt.nDb > 0 && pd_bytes <= pd_limit;
    bool hv_smem = TEAM != 64;  // 64-thread teams: residency matters more than the Hessian's latency (see choose_team)
    if (const char* e = getenv("BMPC_HV_SMEM")) hv_smem = atoi(e) != 0;  // tuning override
    for (int attempt = 0; attempt < 3; ++attempt) {
        layout_smem(h, in_smem, hv_smem);
        h->smem_bytes = h->sm.total * 8 * h->teams_per_cta;
        if (h->smem_bytes <= max_optin) break;
        if (in_smem) { in_smem = false; continue; }
        if (hv_smem) { hv_smem = false; continue; }
        return fail(BMPC_ERR_UNSUPPORTED, "problem too large for shared memory (%d B per CTA)", h->smem_bytes);
    }
    h->pd_in_smem = in_smem;
    h->hv_in_smem = hv_smem;
    CK(bmpc_host::raise_dyn_smem(reinterpret_cast<const void*>(bmpc::step_kernel<TEAM>), h->smem_bytes));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bmpc::step_kernel<TEAM>, CTA, h->smem_bytes));
    if (occ < 1) return fail(BMPC_ERR_UNSUPPORTED, "step kernel does not fit on an SM (smem %d B)", h->smem_bytes);
    const int need = (h->d.N + h->teams_per_cta - 1) / h->teams_per_cta;
    h->grid = std::max(1, std::min(need, occ * h->num_sms));
    return BMPC_OK;
}

const std::vector<bmpc::WarpEntry>& warp_registry() {
    static std::vector<bmpc::WarpEntry> reg = [] {
        std::vector<bmpc::WarpEntry> v;
        bmpc::warp_register_03(v);
        bmpc::warp_register_05(v);
        bmpc::warp_register_07(v);
        bmpc::warp_register_09(v);
        bmpc::warp_register_11(v);
        bmpc::warp_register_13(v);
        bmpc::warp_register_16(v);
        return v;
    }();
    return reg;
}

// warp-per-controller kernel: unit rows (hard bounds on one variable) handled lane-locally, every other row dense;
// one TMA copy per instance, DMMA Hessian build
int configure_warp(bmpc_handle* h, const bmpc::WarpEntry& E) {
    const int NT = E.nt, MP = 32 * E.rpl, nY = h->nY, nx = h->d.nxhat, nS = h->rt.nS;
    const int m = h->rt.nS + h->rt.nDr;  // without the (redundant) eps >= 0 row
    // ---- positions: dense rows first (original order), then the unit rows ----
    std::vector<int> upos(32, MP);       // [0..15] max-side, [16..31] min-side unit row of each variable
    std::vector<char> is_unit(std::max(m, 1), 0);
    for (int r = 0; r < nS; ++r) {
        const int i1 = h->hs_i1[r];
        if (h->hs_i2[r] >= 0 || h->hcc[r] != 0.0 || i1 >= 16) continue;  // 2-variable or soft rows stay dense
        const int side = h->hsig[r] > 0 ? 0 : 16;
        if (upos[side + i1] != MP) continue;  // a second row of the same kind stays dense
        upos[side + i1] = -2 - r;             // marked; the position follows below
        is_unit[r] = 1;
    }
    std::vector<int> pos_row;  // dense positions -> row
    for (int r = 0; r < m; ++r)
        if (!is_unit[r]) pos_row.push_back(r);
    const int mD = (int)pos_row.size();
    const int mh = ((mD + 1) / 2 + 3) & ~3, KS = (mD + 3) / 4;
    const int GR = std::max(2 * mh, 4 * KS);
    if (GR > MP) return BMPC_ERR_UNSUPPORTED;
    std::vector<int4> pinfo(MP, make_int4(-1, 0, -1, 0));
    auto dense_info = [&](int r) {
        int4 pi = make_int4(r, bmpc::PI_VALID, -1, 0);
        if (h->hsig[r] < 0) pi.y |= bmpc::PI_NEG;
        if (r < nS) {
            pi.y |= bmpc::PI_SPARSE;
            pi.z = h->hs_ch[r];
        } else {
            pi.z = h->hdr_src[r - nS];
            if (h->pd_is_ev) {
                const int kb = h->hdr_base[r - nS];
                const int rmax = h->hdb_rmax[kb];
                if (rmax == r || (rmax < 0 && h->hdb_rmin[kb] == r)) pi.y |= bmpc::PI_ISQ;
            }
        }
        return pi;
    };
    for (int p = 0; p < mD; ++p) pinfo[p] = dense_info(pos_row[p]);
    int pu = mD;
    for (int r = 0; r < nS; ++r) {
        if (!is_unit[r]) continue;
        int4 pi = dense_info(r);
        pi.y |= bmpc::PI_UNIT;
        pi.w = h->hs_i1[r];
        pinfo[pu] = pi;
        upos[(h->hsig[r] > 0 ? 0 : 16) + h->hs_i1[r]] = pu;
        ++pu;
    }
    bmpc::WarpLayout& L = h->wp.L;
    int o = 0;
    auto take = [&](int cnt) { int at = o; o += even(std::max(cnt, 1)); return at; };
    // (the arrays up to G sit at the compile-time offsets of WarpSmem<NT, RPL>, bmpc_warp.cuh)
    L.H = take(NT * E.ldh);
    L.L = take(NT * E.ldh);
    L.phi = take(std::max((8 * ((NT + 7) / 8) + 1) * E.ldp, 2 * NT * E.ldn));  // DMMA C tiles + the rhs row, then the columns of L and of M
    L.vx = take(16);
    L.vy = take(16);
    take(16);             // reciprocal pivots (WarpSmem::DV)
    L.w1 = take(MP + 2);  // (+ the "no unit row" slot that reads as zero)
    L.w2 = take(MP + 2);
    L.wd = take(MP + 2);
    L.G = take(GR * E.ldg);
    if (L.G != E.off_g) return fail(BMPC_ERR_CUDA, "warp kernel shared-memory layout mismatch (G at %d, kernel expects %d)", L.G, E.off_g);
    // F and M*Cy are dead before the interior-point loop first writes wd / w2: share the storage when they fit
    if (nY <= MP) {
        L.F = L.wd;
        L.tY = L.w2;
    } else {
        L.F = take(nY);
        L.tY = take(nY);
    }
    L.fx = take(nx);
    L.xh = take(nx);
    L.lu = take(h->d.nu);
    L.dd = take(h->d.nd);
    L.Dh = take(h->d.nd * h->d.Hp);
    L.bar = take(2);
    L.ev = take(h->d.ny);
    L.Fw = take(h->nFw);
    L.total = o;
    if (L.H + NT * E.ldh != L.L) return BMPC_ERR_UNSUPPORTED;  // H and its factor are one TMA copy
    h->smem_bytes = L.total * 8;
    int max_optin = 0;
    CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->d.device));
    if (h->smem_bytes > max_optin) return BMPC_ERR_UNSUPPORTED;
    CK(bmpc_host::raise_dyn_smem(reinterpret_cast<const void*>(E.func), h->smem_bytes));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, E.func, 32, h->smem_bytes));
    if (occ < 1) return BMPC_ERR_UNSUPPORTED;
    if (const char* e = getenv("BMPC_WARP_OCC")) occ = std::max(1, std::min(occ, atoi(e)));  // tuning: resident warps per SM
    h->team = 32;
    h->teams_per_cta = 1;
    h->grid = std::max(1, std::min(h->d.N, occ * h->num_sms));
    h->warp = &E;
    h->pd_in_smem = true;
    const long NM = h->NM;
    const long sG = (long)std::max(GR, 1) * E.ldg, sHL = 2L * NT * E.ldh;
    CK(h->Gw.alloc((size_t)NM * sG));
    CK(h->HLw.alloc((size_t)NM * sHL));
    CK(h->w_pinfo.upload(pinfo, h->stream));
    CK(h->w_upos.upload(upos, h->stream));
    if (pos_row.empty()) pos_row.push_back(0);
    CK(h->w_posrow.upload(pos_row, h->stream));
    const double* pdsrc = h->pd_is_ev ? h->Ev.p : h->Pd.p;
    const long spd = h->pd_is_ev ? h->nEv2 : h->nPd2;
    bmpc::k_make_warp<<<(unsigned)NM, 128, 0, h->stream>>>(pdsrc, spd, h->rt.nDb, h->Hv.p, h->Lv.p, h->nHp2, h->Hee.p, h->rt,
                                                         h->w_posrow.p, mD, h->nz, h->d.neps, NT, E.ldg, E.ldh, GR, h->Gw.p, sG,
                                                         h->HLw.p, sHL);
    h->launches++;
    CK(cudaGetLastError());
    CK(h->lam_ws.alloc((size_t)h->d.N * even(std::max(h->rt.m, 1))));
    CK(h->ws_flag.alloc((size_t)h->d.N));
    CK(cudaMemsetAsync(h->ws_flag.p, 0, (size_t)h->d.N * sizeof(int), h->stream));
    CK(h->order[0].alloc((size_t)h->d.N));
    CK(h->order[1].alloc((size_t)h->d.N));
    CK(h->ocnt.alloc(2));
    CK(cudaMemsetAsync(h->ocnt.p, 0, 2 * sizeof(unsigned int), h->stream));
    h->order_valid = false;
    if (const char* e = getenv("BMPC_WARM")) h->warm_start = atoi(e);
    h->wp.Gw = h->Gw.p;
    h->wp.sGw = h->d.shared_model ? 0 : sG;
    h->wp.HL = h->HLw.p;
    h->wp.sHL = h->d.shared_model ? 0 : sHL;
    h->wp.pinfo = h->w_pinfo.p;
    h->wp.upos = h->w_upos.p;
    h->wp.m = m;
    h->wp.mD = mD;
    h->wp.GR = GR;
#ifdef BMPC_PHASE_CLK
    CK(h->clk.alloc(32));
    CK(cudaMemsetAsync(h->clk.p, 0, 32 * sizeof(long long), h->stream));
    h->wp.clk = h->clk.p;
#endif
    // Next launch's order.  Every instance is filed from the BACK of the list as it completes, so the next launch starts
    // the instances in reverse completion order -- the ones that finished last (the long solves, which stay long from one
    // period to the next) start first: longest-processing-time-first scheduling without a sort.  BMPC_LONG = t restores
    // the two-class variant (>= t iterations to the front, the rest to the back); measured on C1 with the round-2 kernel
    // (tools/studies/env_sweep.sh, same box): t = 12 0.1418-0.1429 ms, 14 0.1404, 18 0.1398, none (this default) 0.1397
    h->wp.long_thresh = 1 << 30;
    if (const char* e = getenv("BMPC_LONG")) h->wp.long_thresh = atoi(e);
    CK(cudaStreamSynchronize(h->stream));  // (the host vectors uploaded above go out of scope)
    return BMPC_OK;
}

int configure_launch(bmpc_handle* h) {
    h->warp = nullptr;
    const int forced = h->d.team ? h->d.team : (getenv("BMPC_TEAM") ? atoi(getenv("BMPC_TEAM")) : 0);
    const int m_rows = h->rt.nS + h->rt.nDr;
    if (forced == 0 && h->n <= 16 && m_rows <= 128 && !h->M_dense && !h->L_dense && h->have_predmat &&
        !(getenv("BMPC_NO_WARP") && atoi(getenv("BMPC_NO_WARP")))) {
        const bmpc::WarpEntry* best = nullptr;
        for (const bmpc::WarpEntry& E : warp_registry()) {
            if (E.nt < h->n || 32 * E.rpl < m_rows) continue;
            if (!best || E.nt < best->nt || (E.nt == best->nt && E.rpl < best->rpl)) best = &E;
        }
        if (best) {
            int rc = configure_warp(h, *best);
            if (rc == BMPC_OK) return rc;
            if (rc != BMPC_ERR_UNSUPPORTED) return rc;
            h->warp = nullptr;
        }
    }
    switch (choose_team(h)) {
        case 8: return configure<8>(h);
        case 16: return configure<16>(h);
        case 32: return configure<32>(h);
        case 64: return configure<64>(h);
        case 128: return configure<128>(h);
        case 256: return configure<256>(h);
        default: return fail(BMPC_ERR_ARG, "team must be one of 0,8,16,32,64,128,256");
    }
}

// (re)build everything that depends on model + weights + constraints together
int finalize(bmpc_handle* h) {
    cudaStream_t s = h->stream;
    if (!h->pd_is_ev && h->rt.nDb > 0) {
        const long tot = (long)h->NM * h->rt.nDb * h->nz;
        if (h->nw > 0) {  // relaxW: the custom rows' matrix, rebuilt from the current Ev
            const long totw = (long)h->NM * h->nFw * h->nz;
            CK(h->Ewv.alloc((size_t)totw));
            bmpc::k_build_ew<<<(unsigned)((totw + 255) / 256), 256, 0, s>>>(h->Ev.p, (long)h->nEv2, h->Wc.p, h->sWc, h->t_blk.p,
                                                                          h->Ewv.p, h->nw, h->d.ny, h->d.nu, h->d.nd, h->d.Hp,
                                                                          h->nY, h->nz, totw);
            h->launches++;
            CK(cudaGetLastError());
        }
        bmpc::k_gather_pd<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(h->Ev.p, h->has_terminal_mats ? h->exv.p : nullptr,
                                                                      h->nw > 0 ? h->Ewv.p : nullptr, h->Pd.p, h->t_pdsrc.p,
                                                                      h->nY, h->d.nxhat, h->nFw, h->nz, h->rt.nDb,
                                                                      (long)h->nEv2, h->nPd2, tot);
        h->launches++;
        CK(cudaGetLastError());
    }
    int rc = configure_launch(h);
    if (rc != BMPC_OK) return rc;
    if (!h->warp) {  // the general kernel keeps the multipliers of the previous period too (IPM warm start)
        CK(h->lam_ws.alloc((size_t)h->d.N * even(std::max(h->rt.m, 1))));
        CK(h->ws_flag.alloc((size_t)h->d.N));
        CK(cudaMemsetAsync(h->ws_flag.p, 0, (size_t)h->d.N * sizeof(int), h->stream));
        if (const char* e = getenv("BMPC_WARM")) h->warm_start = atoi(e);
    }
    CK(cudaStreamSynchronize(s));
    h->dirty = false;
    return BMPC_OK;
}

template <int TEAM>
cudaError_t launch_step(bmpc_handle* h, const bmpc::StepParams& P) {
    constexpr int CTA = bmpc::CtaThreads<TEAM>::value;
    bmpc::step_kernel<TEAM><<<h->grid, CTA, h->smem_bytes, h->stream>>>(P);
    return cudaGetLastError();
}

// copy N*len doubles from a host array (or replicate nothing): returns device pointer via buf
cudaError_t up(DevBuf<double>& buf, const double* src, size_t count, cudaStream_t s) { return buf.upload(src, count, s); }

}  // namespace

extern "C" {

const char* bmpc_last_error(void) { return bmpc_host::last_error().c_str(); }
int bmpc_version(void) { return 100; }

int bmpc_create(bmpc_handle** out, const bmpc_dims* dims, const int32_t* nb) {
    if (!out || !dims || !nb) return fail(BMPC_ERR_ARG, "null argument");
    const bmpc_dims& d = *dims;
    if (d.N < 1 || d.nu < 1 || d.ny < 1 || d.nd < 0 || d.nxhat < 1 || d.Hp < 1 || d.Hc < 1 || d.Hc > d.Hp ||
        (d.neps != 0 && d.neps != 1))
        return fail(BMPC_ERR_ARG, "invalid dimensions");
    int sum = 0;
    for (int i = 0; i < d.Hc; ++i) {
        if (nb[i] < 1) return fail(BMPC_ERR_ARG, "move blocking entries must be >= 1");
        sum += nb[i];
    }
    if (sum != d.Hp) return fail(BMPC_ERR_ARG, "sum(nb) = %d must equal Hp = %d (move_blocking, construct.jl:629-660)", sum, d.Hp);
    if (d.nu * d.Hc + d.neps > 30000) return fail(BMPC_ERR_UNSUPPORTED, "too many decision variables");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(BMPC_ERR_CUDA, "no CUDA device: libbmpc has no CPU fallback (%s)", cudaGetErrorString(e));
    if (d.device < 0 || d.device >= ndev) return fail(BMPC_ERR_ARG, "device %d out of range", d.device);
    CK(cudaSetDevice(d.device));
    bmpc_handle* h = new bmpc_handle();
    h->d = d;
    if (h->d.max_iter <= 0) h->d.max_iter = 50;
    if (!(h->d.tol > 0)) h->d.tol = 1e-11;
    if (const char* e = getenv("BMPC_TOL")) h->d.tol = atof(e);  // diagnostic override (tools/studies)
    h->nb.assign(nb, nb + d.Hc);
    h->blk_start.assign(d.Hc + 1, 0);
    for (int l = 0; l < d.Hc; ++l) {
        h->blk_start[l + 1] = h->blk_start[l] + nb[l];
        for (int t = 0; t < nb[l]; ++t) h->blk_of_t.push_back(l);
    }
    h->nz = d.nu * d.Hc;
    h->n = h->nz + d.neps;
    h->nY = d.ny * d.Hp;
    h->nU = d.nu * d.Hp;
    h->nHp2 = even(h->nz * (h->nz + 1) / 2);
    h->nEv2 = even(h->nY * h->nz);
    h->NM = d.shared_model ? 1 : d.N;
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, d.device);
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return fail(BMPC_ERR_CUDA, "cudaStreamCreate failed");
    }
    h->stream = h->own_stream;
    const size_t N = d.N;
    cudaError_t a = cudaSuccess;
    auto A = [&](cudaError_t r) { if (a == cudaSuccess) a = r; };
    A(h->xhat0.alloc(N * d.nxhat));
    A(h->lastu0.alloc(N * d.nu));
    A(h->ry.alloc(N * d.ny));
    A(h->Z.alloc(N * h->n));
    A(h->u.alloc(N * d.nu));
    A(h->Jv.alloc(N));
    A(h->F.alloc(N * h->nY));
    A(h->qt.alloc(N * h->n));
    A(h->r.alloc(N));
    A(h->lastu_prev.alloc(N * d.nu));
    A(h->status.alloc(N));
    A(h->iters.alloc(N));
    A(h->counters.alloc(2));
    A(h->t_blk.upload(h->blk_of_t, h->stream));
    A(h->t_blkstart.upload(h->blk_start, h->stream));
    if (a == cudaSuccess) a = cudaMemsetAsync(h->counters.p, 0, 2 * sizeof(unsigned), h->stream);
    if (a == cudaSuccess) a = cudaMemsetAsync(h->Z.p, 0, N * h->n * sizeof(double), h->stream);
    if (a == cudaSuccess) a = cudaMemsetAsync(h->lastu0.p, 0, N * d.nu * sizeof(double), h->stream);
    if (a == cudaSuccess) a = cudaStreamSynchronize(h->stream);
    if (a != cudaSuccess) {
        bmpc_destroy(h);
        return fail(BMPC_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(a));
    }
    *out = h;
    return BMPC_OK;
}

int bmpc_destroy(bmpc_handle* h) {
    if (!h) return BMPC_OK;
    cudaSetDevice(h->d.device);
    cudaDeviceSynchronize();
    // every DevBuf member frees its allocation in its destructor (delete h below)
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->zg_ev) cudaEventDestroy(h->zg_ev);
    delete h;
    return BMPC_OK;
}

int bmpc_set_stream(bmpc_handle* h, void* stream) {
    if (!h) return fail(BMPC_ERR_ARG, "null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return BMPC_OK;
}

int bmpc_set_predmat(bmpc_handle* h, const double* E, const double* K, const double* V, const double* B,
                     const double* G, const double* J, const double* Htilde, const double* ex, const double* kx,
                     const double* vx, const double* bx, const double* gx, const double* jx) {
    if (!h || !E || !K || !V || !B || !Htilde) return fail(BMPC_ERR_ARG, "null argument");
    const bmpc_dims& d = h->d;
    if (d.nd > 0 && (!G || !J)) return fail(BMPC_ERR_ARG, "G and J are required when nd > 0");
    const bool term = ex && kx && vx && bx;
    if (term && d.nd > 0 && (!gx || !jx)) return fail(BMPC_ERR_ARG, "gx and jx are required when nd > 0");
    if (h->have_constraints && h->has_terminal_rows && !term)
        return fail(BMPC_ERR_STATE, "terminal constraints are active: the terminal matrices are required");
    CK(cudaSetDevice(d.device));
    cudaStream_t s = h->stream;
    const size_t NM = h->NM, nY = h->nY, nz = h->nz, n = h->n, nx = d.nxhat, nu = d.nu, nd = d.nd, Hp = d.Hp;
    h->route_a = false;
    CK(up(h->E, E, NM * nY * nz, s));
    CK(up(h->K, K, NM * nY * nx, s));
    CK(up(h->V, V, NM * nY * nu, s));
    CK(up(h->B, B, NM * nY, s));
    CK(up(h->Ht, Htilde, NM * n * n, s));
    if (nd > 0) {
        CK(up(h->G, G, NM * nY * nd, s));
        CK(up(h->J, J, NM * nY * nd * Hp, s));
    }
    h->has_terminal_mats = term;
    if (term) {
        CK(up(h->ex, ex, NM * nx * nz, s));
        CK(up(h->kx, kx, NM * nx * nx, s));
        CK(up(h->vx, vx, NM * nx * nu, s));
        CK(up(h->bx, bx, NM * nx, s));
        if (nd > 0) {
            CK(up(h->gx, gx, NM * nx * nd, s));
            CK(up(h->jx, jx, NM * nx * nd * Hp, s));
        }
    }
    // level coordinates: Ev = E*D, exv = ex*D, Hv = D'H̃D (packed), Lv = chol(Hv)
    CK(h->Ev.alloc(NM * h->nEv2));
    CK(cudaMemsetAsync(h->Ev.p, 0, NM * h->nEv2 * sizeof(double), s));
    CK(h->Hv.alloc(NM * h->nHp2));
    CK(h->Lv.alloc(NM * h->nHp2));
    CK(h->Hee.alloc(NM));
    CK(h->lv_ok.alloc(NM));
    CK(cudaMemsetAsync(h->Hv.p, 0, NM * h->nHp2 * sizeof(double), s));
    CK(cudaMemsetAsync(h->Lv.p, 0, NM * h->nHp2 * sizeof(double), s));
    const int TB = 256;
    {
        const long tot = (long)NM * nY * nz;
        bmpc::k_level_cols<<<(unsigned)((tot + TB - 1) / TB), TB, 0, s>>>(h->E.p, h->Ev.p, (int)nY, (int)nz, (int)nu, (long)h->nEv2, tot);
        h->launches++;
    }
    if (term) {
        CK(h->exv.alloc(NM * nx * nz));
        const long tot = (long)NM * nx * nz;
        bmpc::k_level_cols<<<(unsigned)((tot + TB - 1) / TB), TB, 0, s>>>(h->ex.p, h->exv.p, (int)nx, (int)nz, (int)nu, (long)(nx * nz), tot);
        h->launches++;
    }
    {
        const int npair = (int)(nz * (nz + 1) / 2);
        const long tot = (long)NM * npair;
        bmpc::k_level_hess<<<(unsigned)((tot + TB - 1) / TB), TB, 0, s>>>(h->Ht.p, h->Hv.p, h->Hee.p, (int)n, (int)nz, (int)nu,
                                                                         d.neps, h->nHp2, npair, tot);
        bmpc::k_chol_serial<<<(unsigned)((NM + 63) / 64), 64, 0, s>>>(h->Hv.p, h->Lv.p, h->lv_ok.p, (int)nz, h->nHp2, (int)NM);
        h->launches += 2;
    }
    CK(cudaGetLastError());
    h->have_predmat = true;
    h->dirty = true;
    CK(cudaStreamSynchronize(s));
    h->E.release();  // only the level-coordinate copies are kept
    h->ex.release();
    h->Ht.release();
    return BMPC_OK;
}

int bmpc_set_weights(bmpc_handle* h, const double* M, int32_t M_dense, const double* L_diag) {
    return bmpc_set_weights_dense(h, M, M_dense, L_diag, 0);
}

int bmpc_set_weights_dense(bmpc_handle* h, const double* M, int32_t M_dense, const double* L, int32_t L_dense) {
    if (!h || !M) return fail(BMPC_ERR_ARG, "null argument");
    if (h->route_a)
        return fail(BMPC_ERR_STATE, "the weights of a route-A handle are arguments of bmpc_set_model (its Hessian was built from them)");
    CK(cudaSetDevice(h->d.device));
    const size_t NM = h->NM, nY = h->nY, nU = h->nU;
    CK(up(h->Mw, M, NM * (M_dense ? nY * nY : nY), h->stream));
    h->M_dense = M_dense != 0;
    h->has_L = false;
    h->L_dense = false;
    if (L) {
        const size_t cnt = NM * (L_dense ? nU * nU : nU);
        bool any = false;
        for (size_t i = 0; i < cnt && !any; ++i) any = L[i] != 0.0;
        if (any) {  // iszero_L_Hp short-circuit, construct.jl:86 / execute.jl:268
            CK(up(h->Lw, L, cnt, h->stream));
            h->has_L = true;
            h->L_dense = L_dense != 0;
        }
    }
    CK(cudaStreamSynchronize(h->stream));
    h->have_weights = true;
    h->dirty = true;
    return BMPC_OK;
}

int bmpc_set_oppoints(bmpc_handle* h, const double* uop, const double* yop) {
    if (!h) return fail(BMPC_ERR_ARG, "null handle");
    CK(cudaSetDevice(h->d.device));
    const size_t NM = h->NM;
    std::vector<double> z;
    if (uop) {
        CK(up(h->uop, uop, NM * h->d.nu, h->stream));
    } else {
        z.assign(NM * h->d.nu, 0.0);
        CK(h->uop.upload(z, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (yop) {
        CK(up(h->yop, yop, NM * h->d.ny, h->stream));
    } else {
        z.assign(NM * h->d.ny, 0.0);
        CK(h->yop.upload(z, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return BMPC_OK;
}

int bmpc_set_constraints(bmpc_handle* h, const double* U0min, const double* U0max, const double* DUmin,
                         const double* DUmax, const double* Y0min, const double* Y0max, const double* x0min,
                         const double* x0max, const bmpc_softness* soft) {
    if (!h) return fail(BMPC_ERR_ARG, "null handle");
    const bmpc_dims& d = h->d;
    CK(cudaSetDevice(d.device));
    const int N = d.N, nu = d.nu, nz = h->nz, nY = h->nY, nU = h->nU, nx = d.nxhat, neps = d.neps;
    const double INF = INFINITY;
    auto get = [&](const double* a, int len, int i, int k, double dflt) { return a ? a[(size_t)i * len + k] : dflt; };
    auto softv = [&](const double* a, int k, double dflt) { return (neps && a) ? a[k] : (neps ? dflt : 0.0); };
    const bmpc_softness S = soft ? *soft : bmpc_softness{};
    // ---- finiteness pattern (shared by all instances) ----
    const int nFw = h->nFw;
    const int plen = 2 * nU + 2 * nz + 2 * nY + 2 * nx + 2 * nFw;
    std::vector<unsigned char> pat(plen, 0);
    const double* arrs[10] = {U0min, U0max, DUmin, DUmax, Y0min, Y0max, x0min, x0max,
                              h->hWmin.empty() ? nullptr : h->hWmin.data(), h->hWmax.empty() ? nullptr : h->hWmax.data()};
    const int lens[10] = {nU, nU, nz, nz, nY, nY, nx, nx, nFw, nFw};
    for (int i = 0; i < N; ++i) {
        int o = 0;
        for (int a = 0; a < 10; ++a) {
            for (int k = 0; k < lens[a]; ++k, ++o) {
                const double v = get(arrs[a], lens[a], i, k, (a & 1) ? INF : -INF);
                if (std::isnan(v)) return fail(BMPC_ERR_ARG, "NaN bound");
                const unsigned char fin = std::isfinite(v) ? 1 : 0;
                if (i == 0)
                    pat[o] = fin;
                else if (pat[o] != fin)
                    return fail(BMPC_ERR_ARG, "all instances of a handle must share the +-Inf pattern of their bounds");
            }
        }
    }
    if (h->stepped && pat != h->pattern)
        return fail(BMPC_ERR_STATE, "Cannot modify +-Inf constraints after the first step (construct.jl:548-551)");
    const bool any_x = std::any_of(pat.begin() + 2 * nU + 2 * nz + 2 * nY, pat.begin() + 2 * nU + 2 * nz + 2 * nY + 2 * nx,
                                   [](unsigned char c) { return c; });
    const bool any_w = std::any_of(pat.begin() + 2 * nU + 2 * nz + 2 * nY + 2 * nx, pat.end(), [](unsigned char c) { return c; });
    if (any_x && h->have_predmat && !h->has_terminal_mats)
        return fail(BMPC_ERR_ARG, "terminal bounds need the terminal matrices (bmpc_set_predmat ex,kx,vx,bx)");
    // ---- sparse (1-/2-variable) rows, merged by (i1, i2, side, softness, shift channel) ----
    struct Src { int arr, k; };
    typedef std::tuple<int, int, int, double, int> Key;
    std::map<Key, int> groups;
    std::vector<Key> gkeys;
    std::vector<std::vector<Src>> gsrc;
    auto add = [&](int i1, int i2, int side, double c, int ch, int arr, int k) {
        Key key(i1, i2, side, c, ch);
        auto it = groups.find(key);
        int g;
        if (it == groups.end()) {
            g = (int)gkeys.size();
            groups[key] = g;
            gkeys.push_back(key);
            gsrc.emplace_back();
        } else {
            g = it->second;
        }
        gsrc[g].push_back({arr, k});
    };
    for (int t = 0; t < d.Hp; ++t)
        for (int ch = 0; ch < nu; ++ch) {
            const int k = t * nu + ch, var = h->blk_of_t[t] * nu + ch;
            if (pat[k]) add(var, -1, -1, softv(S.C_umin, k, 0.0), ch, 0, k);
            if (pat[nU + k]) add(var, -1, +1, softv(S.C_umax, k, 0.0), ch, 1, k);
        }
    for (int k = 0; k < nz; ++k) {
        const int i2 = k >= nu ? k - nu : -1;
        if (pat[2 * nU + k]) add(k, i2, -1, softv(S.C_dumin, k, 0.0), -1, 2, k);
        if (pat[2 * nU + nz + k]) add(k, i2, +1, softv(S.C_dumax, k, 0.0), -1, 3, k);
    }
    const int nS = (int)gkeys.size();
    // ---- dense rows ----
    std::vector<int> dr_base, dr_src, db_rmax, db_rmin, pd_src;
    std::vector<double> sig, cc;
    std::vector<int> s_i1(nS), s_i2(nS), s_ch(nS);
    for (int g = 0; g < nS; ++g) {
        s_i1[g] = std::get<0>(gkeys[g]);
        s_i2[g] = std::get<1>(gkeys[g]);
        sig.push_back((double)std::get<2>(gkeys[g]));
        cc.push_back(std::get<3>(gkeys[g]));
        s_ch[g] = std::get<4>(gkeys[g]);
    }
    struct DSrc { int arr, k; };
    std::vector<DSrc> dsrc;
    const int oY = 2 * nU + 2 * nz, oX = oY + 2 * nY;
    auto add_dense = [&](int src, bool hasmin, bool hasmax, double cmin, double cmax, int arrmin, int arrmax, int k) {
        if (!hasmin && !hasmax) return;
        const int base = (int)pd_src.size();
        pd_src.push_back(src);
        db_rmax.push_back(-1);
        db_rmin.push_back(-1);
        if (hasmin) {
            db_rmin[base] = nS + (int)dr_base.size();
            dr_base.push_back(base);
            dr_src.push_back(src);
            sig.push_back(-1.0);
            cc.push_back(cmin);
            dsrc.push_back({arrmin, k});
        }
        if (hasmax) {
            db_rmax[base] = nS + (int)dr_base.size();
            dr_base.push_back(base);
            dr_src.push_back(src);
            sig.push_back(+1.0);
            cc.push_back(cmax);
            dsrc.push_back({arrmax, k});
        }
    };
    for (int t = 0; t < nY; ++t)
        add_dense(t, pat[oY + t], pat[oY + nY + t], softv(S.C_ymin, t, 1.0), softv(S.C_ymax, t, 1.0), 4, 5, t);
    for (int i = 0; i < nx; ++i)
        add_dense(nY + i, pat[oX + i], pat[oX + nx + i], softv(S.c_xmin, i, 1.0), softv(S.c_xmax, i, 1.0), 6, 7, i);
    const int oW = oX + 2 * nx;  // custom rows: source nY + nx + r  (bound Wmin/Wmax in absolute units, against Fw[r])
    for (int r = 0; r < nFw; ++r)
        add_dense(nY + nx + r, pat[oW + r], pat[oW + nFw + r], neps ? (h->hCwmin.empty() ? 1.0 : h->hCwmin[r]) : 0.0,
                  neps ? (h->hCwmax.empty() ? 1.0 : h->hCwmax[r]) : 0.0, 8, 9, r);
    const int nDr = (int)dr_base.size(), nDb = (int)pd_src.size();
    // The reference's eps >= 0 row is NOT compiled: every softness weight is non-negative (checked below, as in
    // construct.jl:456-506), so a point with eps < 0 is dominated by the same point with eps = 0 -- the optimum is
    // unchanged, and the row's vanishing multiplier would only slow the interior-point iteration down.
    const int m = nS + nDr;
    if (h->stepped && m != h->rt.m) return fail(BMPC_ERR_STATE, "constraint structure changed after the first step");
    for (double c : cc)
        if (c < 0) return fail(BMPC_ERR_ARG, "softness weights should be non-negative (construct.jl:456)");
    // ---- CSR variable -> sparse rows ----
    std::vector<int> var_ptr(nz + 1, 0), var_row, var_sgn;
    for (int pass = 0; pass < 2; ++pass) {
        std::vector<int> cnt(nz, 0);
        for (int g = 0; g < nS; ++g) {
            const int a = s_i1[g], b = s_i2[g];
            if (pass == 0) {
                var_ptr[a + 1]++;
                if (b >= 0) var_ptr[b + 1]++;
            } else {
                var_row[var_ptr[a] + cnt[a]] = g;
                var_sgn[var_ptr[a] + cnt[a]++] = +1;
                if (b >= 0) {
                    var_row[var_ptr[b] + cnt[b]] = g;
                    var_sgn[var_ptr[b] + cnt[b]++] = -1;
                }
            }
        }
        if (pass == 0) {
            for (int j = 0; j < nz; ++j) var_ptr[j + 1] += var_ptr[j];
            var_row.assign(var_ptr[nz], 0);
            var_sgn.assign(var_ptr[nz], 0);
        }
    }
    std::vector<short> pi, pj;
    for (int i = 0; i < nz; ++i)
        for (int j = 0; j <= i; ++j) {
            pi.push_back((short)i);
            pj.push_back((short)j);
        }
    // ---- per-instance numeric bounds ----
    std::vector<double> sbase((size_t)N * std::max(nS, 1)), dbound((size_t)N * std::max(nDr, 1));
    for (int i = 0; i < N; ++i) {
        for (int g = 0; g < nS; ++g) {
            const double sg = sig[g];
            double best = INF;
            for (const Src& s : gsrc[g]) best = std::min(best, sg * arrs[s.arr][(size_t)i * lens[s.arr] + s.k]);
            sbase[(size_t)i * nS + g] = best;
        }
        for (int r = 0; r < nDr; ++r) dbound[(size_t)i * nDr + r] = arrs[dsrc[r].arr][(size_t)i * lens[dsrc[r].arr] + dsrc[r].k];
    }
    h->hs_i1 = s_i1; h->hs_i2 = s_i2; h->hs_ch = s_ch; h->hdr_src = dr_src; h->hdr_base = dr_base;
    h->hdb_rmax = db_rmax; h->hdb_rmin = db_rmin; h->hsig = sig; h->hcc = cc;
    cudaStream_t s = h->stream;
    CK(h->t_si1.upload(s_i1, s));
    CK(h->t_si2.upload(s_i2, s));
    CK(h->t_sch.upload(s_ch, s));
    CK(h->t_sig.upload(sig, s));
    CK(h->t_c.upload(cc, s));
    CK(h->t_varptr.upload(var_ptr, s));
    CK(h->t_varrow.upload(var_row, s));
    CK(h->t_varsgn.upload(var_sgn, s));
    CK(h->t_dbrmax.upload(db_rmax, s));
    CK(h->t_dbrmin.upload(db_rmin, s));
    CK(h->t_drbase.upload(dr_base, s));
    CK(h->t_drsrc.upload(dr_src, s));
    CK(h->t_pdsrc.upload(pd_src, s));
    CK(h->t_pi.upload(pi, s));
    CK(h->t_pj.upload(pj, s));
    CK(h->sbase.upload(sbase, s));
    CK(h->dbound.upload(dbound, s));
    bmpc::RowTables& rt = h->rt;
    rt.nS = nS;
    rt.nDr = nDr;
    rt.nDb = nDb;
    rt.m = m;
    rt.s_i1 = h->t_si1.p;
    rt.s_i2 = h->t_si2.p;
    rt.s_ch = h->t_sch.p;
    rt.row_sig = h->t_sig.p;
    rt.row_c = h->t_c.p;
    rt.var_ptr = h->t_varptr.p;
    rt.var_row = h->t_varrow.p;
    rt.var_sgn = h->t_varsgn.p;
    rt.db_rmax = h->t_dbrmax.p;
    rt.db_rmin = h->t_dbrmin.p;
    rt.dr_base = h->t_drbase.p;
    rt.dr_src = h->t_drsrc.p;
    rt.pair_i = h->t_pi.p;
    rt.pair_j = h->t_pj.p;
    h->has_terminal_rows = any_x;
    h->pd_is_ev = (nDb == nY) && !any_x && !any_w;  // dense base rows are exactly the rows of Ev, in order
    h->nPd2 = even(nDb * nz);
    if (!h->pd_is_ev && nDb > 0) {
        CK(h->Pd.alloc((size_t)h->NM * h->nPd2));
        CK(cudaMemsetAsync(h->Pd.p, 0, (size_t)h->NM * h->nPd2 * sizeof(double), s));
    }
    CK(cudaStreamSynchronize(s));
    h->pattern = pat;
    h->have_constraints = true;
    h->dirty = true;
    return BMPC_OK;
}

int bmpc_step(bmpc_handle* h, const bmpc_step_io* io) {
    if (!h || !io) return fail(BMPC_ERR_ARG, "null argument");
    if (!h->have_predmat) return fail(BMPC_ERR_STATE, "bmpc_set_predmat / bmpc_set_model must be called before bmpc_step");
    if (!h->have_weights) return fail(BMPC_ERR_STATE, "bmpc_set_weights must be called before bmpc_step");
    const bmpc_dims& d = h->d;
    CK(cudaSetDevice(d.device));
    if (!h->have_constraints) {
        int rc = bmpc_set_constraints(h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc != BMPC_OK) return rc;
    }
    if (!h->uop.p) {
        int rc = bmpc_set_oppoints(h, nullptr, nullptr);
        if (rc != BMPC_OK) return rc;
    }
    const bool resident = io->resident != 0 && io->device_ptrs == 0;
    const bool fused_est = io->xhat0 == nullptr && io->y0m != nullptr;
    if (fused_est && !h->have_estimator) return fail(BMPC_ERR_STATE, "io.y0m needs bmpc_set_estimator");
    if ((!io->xhat0 && !fused_est) || (!io->ry && !io->Rhat_y) || !io->u || !io->status)
        return fail(BMPC_ERR_ARG, "xhat0 (or y0m with a fused estimator), ry|Rhat_y, u, status are required");
    if (!resident && (!io->lastu0 || !io->Ztilde || (!io->iters && !io->host_mapped)))
        return fail(BMPC_ERR_ARG, "lastu0, Ztilde, iters are required unless io.resident = 1");
    if (d.nd > 0 && !io->d0) return fail(BMPC_ERR_ARG, "d0 is required when nd > 0");
    if (h->has_terminal_rows && !h->has_terminal_mats) return fail(BMPC_ERR_STATE, "terminal matrices missing");
    if (h->dirty) {
        int rc = finalize(h);
        if (rc != BMPC_OK) return rc;
    }
    cudaStream_t s = h->stream;
    const size_t N = d.N, nx = d.nxhat, nu = d.nu, ny = d.ny, nd = d.nd, Hp = d.Hp, n = h->n, nY = h->nY, nU = h->nU;
    bmpc::StepParams P{};
    const bool dev = io->device_ptrs != 0;
    // host_mapped: the caller's HOST arrays are page-locked and device-accessible (cudaHostAlloc / cudaHostRegister,
    // unified addressing): the kernel reads the inputs from and writes u/J/status/iters to them directly over PCIe
    // (zero-copy) -- no staging copies.  lastu0 / Z̃ still follow io.resident.
    const bool mapped = !dev && io->host_mapped != 0;
    auto in = [&](DevBuf<double>& buf, const double* src, size_t cnt, const double** dst) -> cudaError_t {
        if (!src) { *dst = nullptr; return cudaSuccess; }
        if (dev || mapped) { *dst = src; return cudaSuccess; }
        cudaError_t e = buf.upload(src, cnt, s);
        *dst = buf.p;
        return e;
    };
    CK(in(h->xhat0, io->xhat0, N * nx, &P.xhat0));
    if (fused_est) {
        CK(in(h->y0m, io->y0m, N * (size_t)h->nym, &P.y0m));
        const long she = d.shared_model ? 0 : 1;
        P.est_on = 1; P.nym = h->nym;
        P.eA = h->eA.p; P.eBu = h->eBu.p; P.eBd = h->eBd.p; P.eCm = h->eCm.p; P.eDdm = h->eDdm.p; P.eK = h->eK.p;
        P.efx = h->e_has_fx ? h->efx.p : nullptr;
        P.s_eA = she * nx * nx; P.s_eBu = she * nx * nu; P.s_eBd = she * nx * nd; P.s_eCm = she * h->nym * nx;
        P.s_eDdm = she * h->nym * nd; P.s_eK = she * nx * h->nym; P.s_efx = she * nx;
        P.xstate = h->xstate.p; P.xcorr = h->xcorr.p;
    }
    CK(in(h->ry, io->ry, N * ny, &P.ry));
    CK(in(h->Rhat_y, io->Rhat_y, N * nY, &P.Rhat_y));
    CK(in(h->Rhat_u, io->Rhat_u, N * nU, &P.Rhat_u));
    CK(in(h->d0, io->d0, N * nd, &P.d0));
    CK(in(h->Dhat0, io->Dhat0, N * nd * Hp, &P.Dhat0));
    CK(in(h->Ys, io->Yhat_s, N * nY, &P.Ys));
    if (io->kkt) {
        if (dev || mapped) {
            P.kkt_out = io->kkt;
        } else {
            CK(h->kkt.alloc(N * 3));
            P.kkt_out = h->kkt.p;
        }
    }
    if (dev) {
        P.lastu0 = io->lastu0;
        P.Z = io->Ztilde;
        P.u = io->u;
        P.J_out = io->J;
        P.status = io->status;
        P.iters = io->iters;
    } else {
        if (!resident) {
            CK(cudaMemcpyAsync(h->lastu0.p, io->lastu0, N * nu * 8, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(h->Z.p, io->Ztilde, N * n * 8, cudaMemcpyHostToDevice, s));
        }
        P.lastu0 = h->lastu0.p;
        P.Z = h->Z.p;
        P.u = mapped ? io->u : h->u.p;
        P.J_out = mapped ? (io->J ? io->J : h->Jv.p) : h->Jv.p;
        P.status = mapped ? io->status : h->status.p;
        P.iters = (mapped && io->iters) ? io->iters : h->iters.p;
    }
    P.N = d.N; P.nu = d.nu; P.ny = d.ny; P.nd = d.nd; P.nx = d.nxhat; P.Hp = d.Hp; P.Hc = d.Hc;
    P.nz = h->nz; P.n = h->n; P.neps = d.neps; P.nY = h->nY; P.nU = h->nU;
    P.max_iter = d.max_iter;
    P.tol = d.tol;
    P.tol_mu = d.tol * 1e-3;  // duality gap  s'lam <= tol_mu * (1+|q|)(1+|h|)
    P.rt = h->rt;
    P.sm = h->sm;
    P.pd_in_smem = h->pd_in_smem; P.pd_is_ev = h->pd_is_ev; P.has_terminal = h->has_terminal_rows;
    P.hv_in_smem = h->hv_in_smem ? 1 : 0;
    P.M_dense = h->M_dense; P.has_L = h->has_L; P.L_dense = h->L_dense ? 1 : 0;
    const long sh = d.shared_model ? 0 : 1;
    P.sEv = sh * h->nEv2; P.sH = sh * h->nHp2; P.sK = sh * nY * nx; P.sV = sh * nY * nu; P.sB = sh * nY;
    P.sG = sh * nY * nd; P.sJ = sh * nY * nd * Hp; P.skx = sh * nx * nx; P.svx = sh * nx * nu; P.sbx = sh * nx;
    P.sgx = sh * nx * nd; P.sjx = sh * nx * nd * Hp; P.sM = sh * (h->M_dense ? nY * nY : nY); P.sL = sh * (h->L_dense ? nU * nU : nU);
    P.suop = sh * nu; P.syop = sh * ny;
    P.Ev = h->Ev.p; P.Hv = h->Hv.p; P.Lv = h->Lv.p; P.Hee = h->Hee.p; P.K = h->K.p; P.V = h->V.p; P.B = h->B.p;
    P.G = h->G.p; P.J = h->J.p; P.kx = h->kx.p; P.vx = h->vx.p; P.bx = h->bx.p; P.gx = h->gx.p; P.jx = h->jx.p;
    P.Mw = h->Mw.p; P.Lw = h->Lw.p; P.uop = h->uop.p; P.yop = h->yop.p; P.sbase = h->sbase.p; P.dbound = h->dbound.p;
    P.lv_ok = h->lv_ok.p; P.blk_of_t = h->t_blk.p; P.blk_start = h->t_blkstart.p;
    if (h->pd_is_ev) { P.Pd = h->Ev.p; P.sPd = P.sEv; } else { P.Pd = h->Pd.p; P.sPd = sh * h->nPd2; }
    P.F_out = h->F.p; P.qt_out = h->qt.p; P.r_out = h->r.p; P.lastu_prev = h->lastu_prev.p;
    P.counters = h->counters.p;
    P.nHp2 = h->nHp2; P.nPd2 = h->nPd2;
    P.lam_ws = h->lam_ws.p; P.ws_flag = h->ws_flag.p; P.ws_stride = even(std::max(h->rt.m, 1));
    P.use_ws = (h->warm_start && h->lam_ws.p) ? 1 : 0;
    P.nw = h->nw; P.sW = d.shared_model ? 0 : h->sWc; P.Wc = h->Wc.p;
#ifdef BMPC_PHASE_CLK
    if (!h->clk.p) {
        CK(h->clk.alloc(32));
        CK(cudaMemsetAsync(h->clk.p, 0, 32 * sizeof(long long), s));
    }
    P.gclk = h->warp ? nullptr : h->clk.p;
#endif
    for (int pr = 0; pr < 8; ++pr) {
        P.zg[pr] = h->zg[pr];
        P.zg_flag[pr] = h->zg_flag[pr];
    }
    P.zg_world = h->zg_world;
    P.zg_rank = h->zg_rank;
    if (h->zg_world > 0) {
        h->zg_epoch++;
        P.zg_epoch = (unsigned long long)h->zg_epoch;
        P.zg_pull = h->zg_pull ? 1 : 0;
        P.zg_timeout = h->zg_timeout.p;
        if (h->zg_pull) {  // this rank's own slot buffer: [slots x N x n]
            P.zg_base = (long)(h->zg_epoch % h->zg_slots) * (long)d.N * (long)n;
            P.zg_need_ack = h->zg_consumer ? (long long)h->zg_epoch - h->zg_slots : 0;  // the period that lived in this slot
        } else {
            P.zg_base = h->zg_row_offset * (long)n;
            P.zg_need_ack = 0;
        }
    }
    cudaError_t le;
    const bool kf = fused_est && h->kf_on;
    int kf_smem = 0;
    if (kf) {  // preparestate! of the time-varying KalmanFilter: gain K̂(k) and corrected covariance, before the step
        const size_t nq = std::max<size_t>(nx, (size_t)h->nym), nym = (size_t)h->nym;
        kf_smem = (int)((2 * nx * nx + 2 * nq * nq + 2 * nx * nym + nym * nym + nq * nq) * 8);
        CK(bmpc_host::raise_dyn_smem(reinterpret_cast<const void*>(bmpc::k_kf_cov), kf_smem));
        bmpc::k_kf_cov<<<(unsigned)h->NM, 64, kf_smem, s>>>((int)h->NM, (int)nx, h->nym, 1, h->eA.p, h->kfQ.p, h->eCm.p, h->kfR.p,
                                                           h->kfP.p, h->eK.p);
        le = cudaGetLastError();
        if (le != cudaSuccess) return fail(BMPC_ERR_CUDA, "KalmanFilter covariance kernel launch failed: %s", cudaGetErrorString(le));
        h->launches++;
    }
    if (h->warp) {
        bmpc::WarpParams Q = h->wp;
        Q.static_first = 1;
        Q.l2_prefetch = 0;  // measured on C1 (L2 flushed before every launch): 0.1744 ms with, 0.1728 ms without -- the launch is
                            // bound by its slowest instance, which starts cold either way; kept as a switch
        if (const char* e = getenv("BMPC_L2_PREFETCH")) Q.l2_prefetch = atoi(e);  // (study override)
        if (const char* e = getenv("BMPC_STATIC_FIRST")) Q.static_first = atoi(e);  // (study override)
        Q.order = h->order_valid ? h->order[h->order_cur].p : nullptr;
        Q.order_next = h->order[h->order_cur ^ 1].p;
        Q.ocnt = h->ocnt.p;
        le = h->warp->launch(P, Q, h->grid, h->smem_bytes, h->stream);
        h->order_cur ^= 1;
        h->order_valid = true;
    } else
    switch (h->team) {
        case 8: le = launch_step<8>(h, P); break;
        case 16: le = launch_step<16>(h, P); break;
        case 32: le = launch_step<32>(h, P); break;
        case 64: le = launch_step<64>(h, P); break;
        case 128: le = launch_step<128>(h, P); break;
        default: le = launch_step<256>(h, P); break;
    }
    if (le != cudaSuccess) return fail(BMPC_ERR_CUDA, "step kernel launch failed: %s", cudaGetErrorString(le));
    h->launches++;
    if (kf) {  // updatestate!: P̂(k+1) = Â P̂(k) Â' + Q̂ (the state prediction ran inside the step kernel)
        bmpc::k_kf_cov<<<(unsigned)h->NM, 64, kf_smem, s>>>((int)h->NM, (int)nx, h->nym, 2, h->eA.p, h->kfQ.p, h->eCm.p, h->kfR.p,
                                                           h->kfP.p, h->eK.p);
        le = cudaGetLastError();
        if (le != cudaSuccess) return fail(BMPC_ERR_CUDA, "KalmanFilter covariance kernel launch failed: %s", cudaGetErrorString(le));
        h->launches++;
    }
    h->stepped = true;
    h->last_Z = P.Z;
    h->last_xhat0 = fused_est ? h->xcorr.p : P.xhat0;
    h->last_d0 = P.d0;
    h->last_Dhat0 = P.Dhat0;
    if (!dev) {
        if (io->lastu0) CK(cudaMemcpyAsync(io->lastu0, h->lastu0.p, N * nu * 8, cudaMemcpyDeviceToHost, s));
        if (io->Ztilde) CK(cudaMemcpyAsync(io->Ztilde, h->Z.p, N * n * 8, cudaMemcpyDeviceToHost, s));
        if (!mapped) {
            CK(cudaMemcpyAsync(io->u, h->u.p, N * nu * 8, cudaMemcpyDeviceToHost, s));
            if (io->J) CK(cudaMemcpyAsync(io->J, h->Jv.p, N * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(io->status, h->status.p, N * 4, cudaMemcpyDeviceToHost, s));
            if (io->iters) CK(cudaMemcpyAsync(io->iters, h->iters.p, N * 4, cudaMemcpyDeviceToHost, s));
            if (io->kkt) CK(cudaMemcpyAsync(io->kkt, h->kkt.p, N * 3 * 8, cudaMemcpyDeviceToHost, s));
        }
    }
    if (io->sync || !dev) CK(cudaStreamSynchronize(s));
    return BMPC_OK;
}

int bmpc_getinfo(bmpc_handle* h, const bmpc_info* info) {
    if (!h || !info) return fail(BMPC_ERR_ARG, "null argument");
    if (!h->stepped) return fail(BMPC_ERR_STATE, "bmpc_getinfo needs a previous bmpc_step");
    const bmpc_dims& d = h->d;
    CK(cudaSetDevice(d.device));
    cudaStream_t s = h->stream;
    const size_t N = d.N, nY = h->nY, nU = h->nU, nx = d.nxhat, n = h->n;
    DevBuf<double> Y, U, X;
    CK(Y.alloc(N * nY));
    CK(U.alloc(N * nU));
    CK(X.alloc(N * nx));
    const long sh = d.shared_model ? 0 : 1;
    bmpc::k_getinfo<<<(unsigned)N, 128, h->nz * sizeof(double), s>>>(
        h->Ev.p, sh * (long)h->nEv2, h->has_terminal_mats ? h->exv.p : nullptr, sh * (long)(nx * h->nz), h->kx.p,
        sh * (long)(nx * nx), h->vx.p, sh * (long)(nx * d.nu), h->bx.p, sh * (long)nx, h->last_Z, h->F.p, h->last_xhat0,
        h->lastu_prev.p, h->t_blk.p, Y.p, U.p, X.p, (int)nY, h->nz, (int)n, d.nu, (int)nx, d.Hp, d.nd, h->gx.p,
        sh * (long)(nx * d.nd), h->jx.p, sh * (long)(nx * d.nd * d.Hp), h->last_d0, h->last_Dhat0);
    h->launches++;
    CK(cudaGetLastError());
    if (info->Yhat0) CK(cudaMemcpyAsync(info->Yhat0, Y.p, N * nY * 8, cudaMemcpyDeviceToHost, s));
    if (info->U0) CK(cudaMemcpyAsync(info->U0, U.p, N * nU * 8, cudaMemcpyDeviceToHost, s));
    if (info->xhat0end && h->has_terminal_mats) CK(cudaMemcpyAsync(info->xhat0end, X.p, N * nx * 8, cudaMemcpyDeviceToHost, s));
    if (info->F) CK(cudaMemcpyAsync(info->F, h->F.p, N * nY * 8, cudaMemcpyDeviceToHost, s));
    if (info->qtilde) CK(cudaMemcpyAsync(info->qtilde, h->qt.p, N * n * 8, cudaMemcpyDeviceToHost, s));
    if (info->r) CK(cudaMemcpyAsync(info->r, h->r.p, N * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    Y.release();
    U.release();
    X.release();
    return BMPC_OK;
}

int bmpc_set_estimator(bmpc_handle* h, const double* Ahat, const double* Buhat, const double* Bdhat, const double* Cmhat,
                       const double* Ddmhat, const double* Khat, const double* fop_minus_xop, int32_t nym) {
    if (!h || !Ahat || !Buhat || !Cmhat) return fail(BMPC_ERR_ARG, "null argument");
    const bmpc_dims& d = h->d;
    if (nym < 1 || nym > d.ny) return fail(BMPC_ERR_ARG, "nym must be 1..ny");
    if (d.nd > 0 && (!Bdhat || !Ddmhat)) return fail(BMPC_ERR_ARG, "Bdhat and Ddmhat are required when nd > 0");
    CK(cudaSetDevice(d.device));
    cudaStream_t s = h->stream;
    const size_t NM = (size_t)h->NM, nx = d.nxhat, nu = d.nu, nd = d.nd, N = d.N;
    CK(h->eA.upload(Ahat, NM * nx * nx, s));
    CK(h->eBu.upload(Buhat, NM * nx * nu, s));
    if (nd) {
        CK(h->eBd.upload(Bdhat, NM * nx * nd, s));
        CK(h->eDdm.upload(Ddmhat, NM * nym * nd, s));
    }
    CK(h->eCm.upload(Cmhat, NM * nym * nx, s));
    if (Khat) {
        CK(h->eK.upload(Khat, NM * nx * nym, s));
    } else {  // the gain will come from the covariance recursion (bmpc_set_estimator_cov)
        CK(h->eK.alloc(NM * nx * nym));
        CK(cudaMemsetAsync(h->eK.p, 0, NM * nx * nym * 8, s));
    }
    h->kf_on = false;
    h->e_has_fx = fop_minus_xop != nullptr;
    if (fop_minus_xop) CK(h->efx.upload(fop_minus_xop, NM * nx, s));
    CK(h->xstate.alloc(N * nx));
    CK(h->xcorr.alloc(N * nx));
    CK(cudaMemsetAsync(h->xstate.p, 0, N * nx * 8, s));
    CK(cudaMemsetAsync(h->xcorr.p, 0, N * nx * 8, s));
    CK(cudaStreamSynchronize(s));
    h->nym = nym;
    h->have_estimator = true;
    return BMPC_OK;
}

int bmpc_set_estimator_cov(bmpc_handle* h, const double* P0, const double* Qhat, const double* Rhat) {
    if (!h || !P0 || !Qhat || !Rhat) return fail(BMPC_ERR_ARG, "null argument");
    if (!h->have_estimator) return fail(BMPC_ERR_STATE, "bmpc_set_estimator_cov needs bmpc_set_estimator");
    if (h->d.nxhat > 32 || h->nym > 32) return fail(BMPC_ERR_UNSUPPORTED, "nxhat <= 32 and nym <= 32 are supported");
    CK(cudaSetDevice(h->d.device));
    cudaStream_t s = h->stream;
    const size_t NM = (size_t)h->NM, nx = h->d.nxhat, nym = h->nym;
    CK(h->kfP.upload(P0, NM * nx * nx, s));
    CK(h->kfQ.upload(Qhat, NM * nx * nx, s));
    CK(h->kfR.upload(Rhat, NM * nym * nym, s));
    CK(cudaStreamSynchronize(s));
    h->kf_on = true;
    return BMPC_OK;
}

int bmpc_get_cov(bmpc_handle* h, double* Phat) {
    if (!h || !Phat) return fail(BMPC_ERR_ARG, "null argument");
    if (!h->kf_on) return fail(BMPC_ERR_STATE, "bmpc_get_cov needs bmpc_set_estimator_cov");
    CK(cudaSetDevice(h->d.device));
    CK(cudaMemcpyAsync(Phat, h->kfP.p, (size_t)h->NM * h->d.nxhat * h->d.nxhat * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BMPC_OK;
}

int bmpc_set_state(bmpc_handle* h, const double* xhat0) {
    if (!h || !xhat0) return fail(BMPC_ERR_ARG, "null argument");
    if (!h->have_estimator) return fail(BMPC_ERR_STATE, "bmpc_set_state needs bmpc_set_estimator");
    CK(cudaSetDevice(h->d.device));
    CK(cudaMemcpyAsync(h->xstate.p, xhat0, (size_t)h->d.N * h->d.nxhat * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BMPC_OK;
}

int bmpc_get_state(bmpc_handle* h, double* xhat0, double* xhat0_corrected) {
    if (!h) return fail(BMPC_ERR_ARG, "null handle");
    if (!h->have_estimator) return fail(BMPC_ERR_STATE, "bmpc_get_state needs bmpc_set_estimator");
    CK(cudaSetDevice(h->d.device));
    const size_t bytes = (size_t)h->d.N * h->d.nxhat * 8;
    if (xhat0) CK(cudaMemcpyAsync(xhat0, h->xstate.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (xhat0_corrected) CK(cudaMemcpyAsync(xhat0_corrected, h->xcorr.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BMPC_OK;
}

int bmpc_set_gather_pull(bmpc_handle* h, void* const* peer_bufs, void* const* peer_flags, int32_t world, int32_t rank,
                         const int32_t* row_offsets, int32_t slots) {
    if (!h) return fail(BMPC_ERR_ARG, "null handle");
    if (world == 0 || !peer_bufs) {
        h->zg_world = 0;
        h->zg_pull = false;
        return BMPC_OK;
    }
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(BMPC_ERR_ARG, "world must be 1..8 and 0 <= rank < world");
    if (!peer_flags || !row_offsets) return fail(BMPC_ERR_ARG, "null argument");
    if (slots < 2) return fail(BMPC_ERR_ARG, "slots must be >= 2 (the period being written and the one being read)");
    for (int pr = 0; pr < world; ++pr)
        if (row_offsets[pr + 1] < row_offsets[pr]) return fail(BMPC_ERR_ARG, "row_offsets must be non-decreasing");
    if (row_offsets[rank + 1] - row_offsets[rank] != h->d.N)
        return fail(BMPC_ERR_ARG, "row_offsets[rank+1] - row_offsets[rank] must equal this handle's N");
    for (int pr = 0; pr < world; ++pr) {
        if (!peer_bufs[pr] || !peer_flags[pr]) return fail(BMPC_ERR_ARG, "null peer buffer");
        h->zg[pr] = static_cast<double*>(peer_bufs[pr]);
        h->zg_flag[pr] = static_cast<unsigned long long*>(peer_flags[pr]);
    }
    for (int pr = world; pr < 8; ++pr) { h->zg[pr] = nullptr; h->zg_flag[pr] = nullptr; }
    for (int pr = 0; pr <= world; ++pr) h->zg_row_off[pr] = row_offsets[pr];
    CK(cudaSetDevice(h->d.device));
    CK(h->zg_timeout.alloc(1));
    CK(h->zg_done.alloc(1));
    CK(cudaMemsetAsync(h->zg_timeout.p, 0, sizeof(int), h->stream));
    CK(cudaMemsetAsync(h->zg_done.p, 0, sizeof(unsigned int), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->zg_world = world;
    h->zg_rank = rank;
    h->zg_slots = slots;
    h->zg_pull = true;
    h->zg_consumer = false;
    h->zg_epoch = 0;
    return BMPC_OK;
}

int bmpc_set_gather(bmpc_handle* h, void* const* peer_bufs, int32_t world, int32_t rank) {
    if (!h) return fail(BMPC_ERR_ARG, "null handle");
    if (world == 0 || !peer_bufs) {
        h->zg_world = 0;
        h->zg_pull = false;
        return BMPC_OK;
    }
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(BMPC_ERR_ARG, "world must be 1..8 and 0 <= rank < world");
    // push protocol: the epilogue stores into row block `rank` of EVERY peer's [world x N x n] buffer (equal shard sizes);
    // the caller supplies the cross-rank barrier before reading
    for (int pr = 0; pr < world; ++pr) {
        if (!peer_bufs[pr]) return fail(BMPC_ERR_ARG, "null peer buffer");
        h->zg[pr] = static_cast<double*>(peer_bufs[pr]);
        h->zg_flag[pr] = nullptr;
    }
    for (int pr = world; pr < 8; ++pr) { h->zg[pr] = nullptr; h->zg_flag[pr] = nullptr; }
    h->zg_world = world;
    h->zg_rank = rank;
    h->zg_row_offset = (long)rank * h->d.N;
    h->zg_slots = 1;
    h->zg_pull = false;
    h->zg_epoch = 0;
    return BMPC_OK;
}

int64_t bmpc_gather_epoch(bmpc_handle* h) { return h ? h->zg_epoch : 0; }

int bmpc_gather_pull(bmpc_handle* h, int64_t epoch, double* dst, void* stream) {
    if (!h) return fail(BMPC_ERR_ARG, "null handle");
    if (h->zg_world <= 0 || !h->zg_pull) return fail(BMPC_ERR_STATE, "bmpc_gather_pull needs bmpc_set_gather_pull");
    if (epoch < 1 || epoch > h->zg_epoch) return fail(BMPC_ERR_ARG, "epoch must be 1..bmpc_gather_epoch()");
    if (epoch + h->zg_slots <= h->zg_epoch + 1 && !h->zg_consumer)
        return fail(BMPC_ERR_STATE, "that period's slot has already been overwritten (pull every period, or use more slots)");
    CK(cudaSetDevice(h->d.device));
    bmpc::PullParams Q{};
    for (int pr = 0; pr < h->zg_world; ++pr) {
        Q.src[pr] = h->zg[pr];
        Q.flags[pr] = h->zg_flag[pr];
    }
    for (int pr = 0; pr <= h->zg_world; ++pr) Q.row_off[pr] = h->zg_row_off[pr];
    Q.world = h->zg_world; Q.rank = h->zg_rank; Q.n = h->n; Q.slots = h->zg_slots; Q.parts = 8;
    Q.epoch = (unsigned long long)epoch;
    Q.limit = 4000000000LL;  // ~2 s at 2 GHz: a peer that never publishes must not hang the device
    Q.timed_out = h->zg_timeout.p;
    Q.done = h->zg_done.p;
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    if (s != h->stream) {
        // a side stream: the pull starts once this rank's own step kernel of `epoch` has finished, so its spinning
        // CTAs never take SM residency away from that launch
        if (!h->zg_ev) CK(cudaEventCreateWithFlags(&h->zg_ev, cudaEventDisableTiming));
        CK(cudaEventRecord(h->zg_ev, h->stream));
        CK(cudaStreamWaitEvent(s, h->zg_ev, 0));
    }
    bmpc::k_gather_pull<<<h->zg_world * Q.parts, 256, 0, s>>>(Q, dst);
    CK(cudaGetLastError());
    h->launches++;
    h->zg_consumer = true;  // from now on a launch waits for every reader's ack before it reuses a slot
    return BMPC_OK;
}

int bmpc_gather_timed_out(bmpc_handle* h) {
    if (!h || !h->zg_timeout.p) return 0;
    int v = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpy(&v, h->zg_timeout.p, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return v;
}

int bmpc_get_states(bmpc_handle* h, double* X0) {
    if (!h || !X0) return fail(BMPC_ERR_ARG, "null argument");
    if (!h->stepped) return fail(BMPC_ERR_STATE, "bmpc_get_states needs a previous bmpc_step");
    if (!h->have_model_mats) return fail(BMPC_ERR_STATE, "bmpc_get_states needs the augmented model (bmpc_set_model, route A)");
    const bmpc_dims& d = h->d;
    CK(cudaSetDevice(d.device));
    cudaStream_t s = h->stream;
    const size_t N = d.N, nx = d.nxhat;
    DevBuf<double> X;
    CK(X.alloc(N * nx * d.Hp));
    const long sh = d.shared_model ? 0 : 1;
    const int threads = 32 * (int)((nx + 31) / 32);
    bmpc::k_ms_states<<<(unsigned)N, threads, (2 * nx + d.nu) * sizeof(double), s>>>(
        h->mA.p, sh * (long)(nx * nx), h->mBu.p, sh * (long)(nx * d.nu), h->mBd.p, sh * (long)(nx * d.nd), h->mf.p, sh * (long)nx,
        h->last_Z, h->last_xhat0, h->lastu_prev.p, h->last_d0, h->last_Dhat0, h->t_blk.p, X.p, (int)nx, d.nu, d.nd, d.Hp, h->n);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(X0, X.p, N * nx * d.Hp * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return BMPC_OK;
}

int bmpc_set_custom(bmpc_handle* h, int32_t nw, const double* Wy, const double* Wu, const double* Wd, const double* Wr,
                    const double* Chat, const double* Ddhat, const double* dop) {
    if (!h) return fail(BMPC_ERR_ARG, "null handle");
    const bmpc_dims& d = h->d;
    if (h->stepped) return fail(BMPC_ERR_STATE, "custom linear constraints cannot change after the first step");
    if (nw < 0) return fail(BMPC_ERR_ARG, "nw must be >= 0");
    if (nw == 0) {
        h->nw = h->nFw = 0;
        h->hWmin.clear(); h->hWmax.clear(); h->hCwmin.clear(); h->hCwmax.clear();
        h->dirty = true;
        return BMPC_OK;
    }
    if (!Chat) return fail(BMPC_ERR_ARG, "Chat (estim.Ĉ) is required: ŷ(k) enters the first block of the custom rows");
    if (d.nd > 0 && !Ddhat) return fail(BMPC_ERR_ARG, "Ddhat is required when nd > 0");
    CK(cudaSetDevice(d.device));
    const size_t NM = (size_t)h->NM, ny = d.ny, nu = d.nu, nd = d.nd, nx = d.nxhat, w = (size_t)nw;
    const size_t per = w * (2 * ny + nu + nd) + ny * nx + ny * nd + nd;
    std::vector<double> pack(NM * per, 0.0);
    for (size_t i = 0; i < NM; ++i) {
        double* o = pack.data() + i * per;
        auto put = [&](const double* src, size_t cnt) {
            if (src) std::copy(src + i * cnt, src + (i + 1) * cnt, o);
            o += cnt;
        };
        put(Wy, w * ny); put(Wu, w * nu); put(Wd, w * nd); put(Wr, w * ny); put(Chat, ny * nx); put(Ddhat, ny * nd); put(dop, nd);
    }
    CK(h->Wc.upload(pack, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->sWc = (long)per;
    h->nw = nw;
    h->nFw = nw * (d.Hp + 1);
    h->hWmin.clear(); h->hWmax.clear(); h->hCwmin.clear(); h->hCwmax.clear();
    h->dirty = true;
    return BMPC_OK;
}

int bmpc_set_custom_bounds(bmpc_handle* h, const double* Wmin, const double* Wmax, const double* C_wmin, const double* C_wmax) {
    if (!h) return fail(BMPC_ERR_ARG, "null handle");
    if (h->nw <= 0) return fail(BMPC_ERR_STATE, "bmpc_set_custom_bounds needs bmpc_set_custom");
    const size_t cnt = (size_t)h->d.N * h->nFw;
    auto chk = [&](const double* a, size_t n) { for (size_t i = 0; a && i < n; ++i) if (std::isnan(a[i])) return false; return true; };
    if (!chk(Wmin, cnt) || !chk(Wmax, cnt)) return fail(BMPC_ERR_ARG, "NaN bound");
    for (int r = 0; r < h->nFw; ++r)
        if ((C_wmin && C_wmin[r] < 0) || (C_wmax && C_wmax[r] < 0)) return fail(BMPC_ERR_ARG, "softness weights should be non-negative");
    if ((C_wmin || C_wmax) && !h->d.neps) return fail(BMPC_ERR_ARG, "Slack variable weight Cwt must be finite to set softness parameters");
    h->hWmin.assign(cnt, -INFINITY);
    h->hWmax.assign(cnt, INFINITY);
    if (Wmin) h->hWmin.assign(Wmin, Wmin + cnt);
    if (Wmax) h->hWmax.assign(Wmax, Wmax + cnt);
    if (C_wmin) h->hCwmin.assign(C_wmin, C_wmin + h->nFw); else h->hCwmin.clear();
    if (C_wmax) h->hCwmax.assign(C_wmax, C_wmax + h->nFw); else h->hCwmax.clear();
    return BMPC_OK;  // compiled into the row tables by the next bmpc_set_constraints
}

int bmpc_launch_info(bmpc_handle* h, int32_t out[8]) {
    if (!h || !out) return fail(BMPC_ERR_ARG, "null argument");
    out[0] = h->team;
    out[1] = h->teams_per_cta;
    out[2] = h->grid;
    out[3] = h->smem_bytes;
    // >= 200: warp-per-controller kernel, NT = value - 200; else 0/1 = dense rows staged in shared memory
    out[4] = h->warp ? 200 + h->warp->nt : (int)h->pd_in_smem;
    out[5] = h->rt.m;
    out[6] = h->rt.nS;
    out[7] = h->rt.nDr;
    return BMPC_OK;
}

int64_t bmpc_launch_count(bmpc_handle* h) { return h ? h->launches : 0; }

#ifdef BMPC_PHASE_CLK
// study builds only (tools/studies/phase_clk.py): read and reset the per-phase cycle accumulators
int bmpc_debug_phase_clk(bmpc_handle* h, long long out[32]) {
    if (!h || !h->clk.p) return BMPC_ERR_STATE;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out, h->clk.p, 32 * sizeof(long long), cudaMemcpyDeviceToHost));
    CK(cudaMemset(h->clk.p, 0, 32 * sizeof(long long)));
    return BMPC_OK;
}
#endif

int bmpc_set_model(bmpc_handle* h, const double* Ahat, const double* Buhat, const double* Chat, const double* Bdhat,
                   const double* Ddhat, const double* fop_minus_xop, const double* M_diag, const double* N_diag,
                   const double* L_diag, double Cwt) {
    if (!h || !Ahat || !Buhat || !Chat || !M_diag || !N_diag) return fail(BMPC_ERR_ARG, "null argument");
    const bmpc_dims& d = h->d;
    if (d.nd > 0 && (!Bdhat || !Ddhat)) return fail(BMPC_ERR_ARG, "Bdhat and Ddhat are required when nd > 0");
    if (d.neps && !(Cwt >= 0 && std::isfinite(Cwt))) return fail(BMPC_ERR_ARG, "Cwt must be finite and >= 0 when neps = 1");
    CK(cudaSetDevice(d.device));
    cudaStream_t s = h->stream;
    const size_t NM = h->NM, nY = h->nY, nz = h->nz, nx = d.nxhat, nu = d.nu, ny = d.ny, nd = d.nd, Hp = d.Hp, nU = h->nU;
    DevBuf<double>&A = h->mA, &Bu = h->mBu, &Bd = h->mBd, &f = h->mf;
    DevBuf<double> C, Dd, Nd;
    std::vector<double> zeros;
    CK(A.upload(Ahat, NM * nx * nx, s));
    CK(Bu.upload(Buhat, NM * nx * nu, s));
    CK(C.upload(Chat, NM * ny * nx, s));
    if (nd) {
        CK(Bd.upload(Bdhat, NM * nx * nd, s));
        CK(Dd.upload(Ddhat, NM * ny * nd, s));
    }
    if (fop_minus_xop) {
        CK(f.upload(fop_minus_xop, NM * nx, s));
    } else {
        zeros.assign(NM * nx, 0.0);
        CK(f.upload(zeros, s));
        CK(cudaStreamSynchronize(s));
    }
    CK(up(h->Mw, M_diag, NM * nY, s));
    CK(Nd.upload(N_diag, NM * nz, s));
    std::vector<double> Lz;
    const double* Lsrc = L_diag;
    h->has_L = false;
    if (L_diag) {
        for (size_t i = 0; i < NM * nU && !h->has_L; ++i) h->has_L = L_diag[i] != 0.0;
    } else {
        Lz.assign(NM * nU, 0.0);
        Lsrc = Lz.data();
    }
    CK(up(h->Lw, Lsrc, NM * nU, s));
    h->M_dense = false;
    CK(h->K.alloc(NM * nY * nx));
    CK(h->V.alloc(NM * nY * nu));
    CK(h->B.alloc(NM * nY));
    if (nd) {
        CK(h->G.alloc(NM * nY * nd));
        CK(h->J.alloc(NM * nY * nd * Hp));
        CK(h->gx.alloc(NM * nx * nd));
        CK(h->jx.alloc(NM * nx * nd * Hp));
    }
    CK(h->kx.alloc(NM * nx * nx));
    CK(h->vx.alloc(NM * nx * nu));
    CK(h->bx.alloc(NM * nx));
    CK(h->Ev.alloc(NM * h->nEv2));
    CK(h->exv.alloc(NM * nx * nz));
    CK(h->Hv.alloc(NM * h->nHp2));
    CK(h->Lv.alloc(NM * h->nHp2));
    CK(h->Hee.alloc(NM));
    CK(h->lv_ok.alloc(NM));
    CK(cudaMemsetAsync(h->Ev.p, 0, NM * h->nEv2 * sizeof(double), s));
    CK(cudaMemsetAsync(h->Hv.p, 0, NM * h->nHp2 * sizeof(double), s));
    CK(cudaMemsetAsync(h->Lv.p, 0, NM * h->nHp2 * sizeof(double), s));
    bmpc::ModelParams P{};
    P.nu = d.nu; P.ny = d.ny; P.nd = d.nd; P.nx = d.nxhat; P.Hp = d.Hp; P.Hc = d.Hc; P.nz = h->nz; P.nY = h->nY;
    P.nU = h->nU; P.neps = d.neps; P.nHp2 = h->nHp2; P.nEv2 = h->nEv2; P.Cwt = Cwt;
    P.A = A.p; P.Bu = Bu.p; P.C = C.p; P.Bd = Bd.p; P.Dd = Dd.p; P.f = f.p; P.Mdiag = h->Mw.p; P.Ndiag = Nd.p;
    P.Ldiag = h->Lw.p;
    P.K = h->K.p; P.V = h->V.p; P.B = h->B.p; P.G = h->G.p; P.J = h->J.p; P.kx = h->kx.p; P.vx = h->vx.p;
    P.bx = h->bx.p; P.gx = h->gx.p; P.jx = h->jx.p; P.Ev = h->Ev.p; P.exv = h->exv.p; P.Hv = h->Hv.p;
    P.Hee = h->Hee.p; P.blk_start = h->t_blkstart.p;
    const size_t smem = (3 * nx * nx + 3 * ny * nx + 3 * nx * nu + 5 * nx + 2 * nx * nd + 8) * sizeof(double);
    if (smem > 200 * 1024) return fail(BMPC_ERR_UNSUPPORTED, "model too large for the on-device builder (use bmpc_set_predmat)");
    CK(bmpc_host::raise_dyn_smem(reinterpret_cast<const void*>(bmpc::k_build_model), (int)smem));
    bmpc::k_build_model<<<(unsigned)NM, 128, smem, s>>>(P);
    bmpc::k_chol_serial<<<(unsigned)((NM + 63) / 64), 64, 0, s>>>(h->Hv.p, h->Lv.p, h->lv_ok.p, (int)nz, h->nHp2, (int)NM);
    h->launches += 2;
    CK(cudaGetLastError());
    h->has_terminal_mats = true;
    h->have_predmat = true;
    h->have_weights = true;
    h->route_a = true;
    h->L_dense = false;
    h->dirty = true;
    CK(cudaStreamSynchronize(s));
    C.release(); Dd.release(); Nd.release();
    h->have_model_mats = true;
    return BMPC_OK;
}

}  // extern "C"
