// Warp-per-controller step kernel for SMALL problems (n = nu*Hc + neps <= 16, m <= 128 rows): the
// shape of BASELINE.json's headline config (C1: n = 11, 60 rows).
//
// One warp owns one controller instance for the whole moveinput!:
//   * rows come in two classes.  UNIT rows -- hard bounds on ONE variable, +-x_i <= h, at most a max- and a min-side
//     row per variable (C1: the 20 merged input-box rows) -- never enter a matrix: their row product is +-x_i, their
//     contribution to G'w is +-w_r on lane i and to Phi a diagonal term.  Every other row (predicted outputs,
//     terminal states, 2-variable increment rows, soft rows) is a DENSE row of Gt = [sigma*p_r, -c_r] (mD x n,
//     row-major, zero padded), loaded per instance by ONE TMA bulk copy (cp.async.bulk + mbarrier) together with the
//     Hessian and its cached Cholesky factor.  Rows are addressed by POSITION p: dense rows first, then the unit rows;
//   * lane l owns positions l, l+32, ... (s, lambda, h, r_p in registers) and, for l < n, decision variable l;
//   * Phi = H + Gt' D Gt is formed on the FP64 TENSOR pipe: mma.sync.m8n8k4.f64 (DMMA) tiles A = Gt' (8 variables x 4
//     rows), B = D*Gt, accumulators initialised with H; only the lower block-triangle is computed;
//   * factorisation Phi = M D M' (LDL', M unit lower): right-looking, one row per lane in registers, column k
//     broadcast through shared memory UNSCALED (its store, the warp barrier and the broadcast loads run while the
//     pivot's reciprocal is in flight); reciprocals by rcp_fast (hardware seed + one third-order step).  The matrix is
//     BORDERED by the predictor's right-hand side (lane NT): the same column updates leave D^-1 M^-1 rhs there, so the
//     predictor's forward substitution costs no extra instruction.  M stays in shared memory (column-major, zeros on
//     and above the diagonal); each substitution sweep re-reads its row / column of M with independent loads and then
//     runs n/2 dependent (shuffle, 2 FMA) steps, two variables at a time.  (Measured and rejected: an explicit
//     W = D^-1 L^-1 with the solves as two mat-vecs, and 2 x 2 pivot blocks with an explicit block inverse -- shorter
//     iterations, but the late, ill-conditioned iterations of the degenerate instances lose accuracy: more
//     iterations, non-optimal exits);
//   * G'w products read Gt columns from shared memory with the row range split between the two half-warps;
//   * the redundant  eps >= 0  row of the reference QP is NOT compiled: every softness weight is
//     non-negative (construct.jl:456-506), so any point with eps < 0 is dominated by the same point with
//     eps = 0 and the optimum is unchanged -- while the row's vanishing multiplier (no strict
//     complementarity whenever no soft constraint is active) is what slows interior-point convergence.
// Algorithm, tolerances, status policy and outputs are those of the general kernel (bmpc_device.cuh),
// with two refinements: the step-to-boundary fraction tends to 1 as the affine step closes the gap
// (tau = max(0.99, 1 - mu_aff/mu)), and work is handed out in the REVERSE of the previous launch's completion order
// (the long solves, which finished last and stay long from one period to the next, start first) so that the few
// 20-iteration solves do not start late in a launch.
#pragma once
#include "bmpc_device.cuh"

namespace bmpc {

struct WarpLayout {  // per-warp shared-memory offsets (doubles)
    int G, H, L, phi, vx, vy, w1, w2, wd, F, tY, fx, xh, lu, dd, Dh, bar, ev, Fw, total;
};

// per-position row descriptor (built on the host, bmpc_api.cu configure_warp)
constexpr int PI_VALID = 1, PI_SPARSE = 2, PI_UNIT = 4, PI_ISQ = 8, PI_NEG = 16;
struct WarpParams {
    WarpLayout L;
    const double* Gw;   // [GR x LDG] per instance: the DENSE rows, zero padded
    long sGw;
    const double* HL;   // [2 x NT x LDH] per instance: extended Hessian, then its factor (1/L_ii on the diagonal)
    long sHL;
    const int4* pinfo;  // [32 RPL] per position: x = row index r in the row tables, y = PI_* flags,
                        //   z = dense: source of the bound (t < nY: F[t], else fx) / sparse: input channel shifting it or -1,
                        //   w = unit rows: the variable
    const int* upos;    // [2 x 16] position of the max-side / min-side unit row of each variable (32 RPL = none)
    const int* order;   // processing order of this launch (nullptr: 0..N-1)
    int* order_next;    // written by this launch: the next launch's order (filed from the back in completion order)
    unsigned int* ocnt; // [2] fill counters of order_next (front, back)
    int long_thresh;    // iterations from which an instance is filed at the front of order_next (default: never)
    int l2_prefetch;    // 1: every CTA prefetches into L2 the constants of the instance one queue wave ahead of its own
    int static_first;   // 1: the first slot of every CTA is its block index (all CTAs resident), 0: every slot from the counter
    int m;              // inequality rows (without the eps >= 0 row)
    int mD;             // dense rows = positions 0..mD-1; the unit rows follow
    int GR;             // rows of Gw (mD rounded up for the half-warp split and the DMMA k-steps)
    long long* clk;     // BMPC_PHASE_CLK study builds only: [32] accumulated cycles per phase (nullptr otherwise)
};

template <int NT>
struct WarpDims {
    static constexpr int NB = (NT + 7) / 8;            // 8-wide variable blocks (DMMA tiles)
    static constexpr int LDG = NT <= 12 ? 12 : 20;     // == 4 (mod 8): conflict-free DMMA fragment loads
    static constexpr int LDH = (NT + 1) & ~1;
    static constexpr int LDN0 = (NT + 1) & ~1;
    static constexpr int LDN = (LDN0 % 4 == 0) ? LDN0 + 2 : LDN0;  // factor columns: even (16-byte rows), == 2 (mod 4): conflict-free LDS.128
    static constexpr int LDP = 8 * NB + 2;
};

// Shared-memory offsets (doubles) of the arrays the interior-point loop touches: compile-time constants, so their
// addresses are immediates instead of a constant-bank load + LEA at every use (the host lays the CTA out in the same
// order, configure_warp in bmpc_api.cu, and checks G against WarpEntry::off_g).  The per-instance sized arrays follow G.
template <int NT, int RPL>
struct WarpSmem {
    using D = WarpDims<NT>;
    static constexpr int even_(int v) { return (v + 1) & ~1; }
    static constexpr int MP = 32 * RPL;
    static constexpr int H = 0;
    static constexpr int L = H + NT * D::LDH;  // (H and its factor are one TMA copy)
    static constexpr int PHI = L + NT * D::LDH;
    static constexpr int PHI_SZ = even_((8 * D::NB + 1) * D::LDP > 2 * NT * D::LDN ? (8 * D::NB + 1) * D::LDP : 2 * NT * D::LDN);
    static constexpr int VX = PHI + PHI_SZ;
    static constexpr int VY = VX + 16;
    static constexpr int DV = VY + 16;  // reciprocal pivots of the factorisation
    static constexpr int W1 = DV + 16;
    static constexpr int W2 = W1 + MP + 2;
    static constexpr int WD = W2 + MP + 2;
    static constexpr int G = WD + MP + 2;
};

constexpr unsigned WFULL = 0xffffffffu;

// Study builds (-DBMPC_PHASE_CLK, tools/studies/phase_clk.py): lane 0 accumulates the cycles spent between marks.
#ifdef BMPC_PHASE_CLK
#define PCLK_DECL long long pclk_t = clock64(); long long pclk_acc[24] = {0}
#define PCLK(i) do { const long long t_ = clock64(); pclk_acc[i] += t_ - pclk_t; pclk_t = t_; } while (0)
#define PCLK_FLUSH() do { if (lane == 0 && Q.clk) { for (int i_ = 0; i_ < 24; ++i_) if (pclk_acc[i_]) atomicAdd((unsigned long long*)&Q.clk[i_], (unsigned long long)pclk_acc[i_]); } for (int i_ = 0; i_ < 24; ++i_) pclk_acc[i_] = 0; } while (0)
#else
#define PCLK_DECL
#define PCLK(i)
#define PCLK_FLUSH()
#endif

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(WFULL, v, o);
    return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(WFULL, v, o));
    return v;
}
__device__ __forceinline__ double wmin(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(WFULL, v, o));
    return v;
}
// 1/d without the rounding guarantees (and the slow path) of __drcp_rn: the hardware seed (MUFU.RCP64H, ~20 bits) and
// ONE third-order step r (1 + e + e^2), e = 1 - d r: three dependent DFMAs instead of four plus fix-up code, relative
// error e^3 < 1e-17 before rounding (a few ulp).  For positive normal d only (0 / denormals give Inf / NaN): every
// caller guards its argument.
__device__ __forceinline__ double rcp_fast(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    return fma(r, fma(e, e, e), r);
}
// three maxima and one sum in ONE butterfly (the shuffles of the four values overlap).  The maxima feed thresholds
// and scale estimates only: they are reduced in fp32 (one shuffle + one FMNMX each, rounded UP so that a convergence
// test can only become stricter); the sum stays fp64.
__device__ __forceinline__ float f32_up(double v) { return __double2float_ru(v); }
// All the maxima are of NON-NEGATIVE values, whose fp32 bit patterns order like unsigned integers: one REDUX
// (redux.sync.max.u32) per maximum replaces a five-step shuffle butterfly (a NaN, whose pattern is above +Inf, wins).
__device__ __forceinline__ float wmax_nonneg_f32(float v) {
    return __uint_as_float(__reduce_max_sync(WFULL, __float_as_uint(v)));
}
__device__ __forceinline__ void wred_mmms(double& a, double& b, double& c, double& s) {
    const float fa = wmax_nonneg_f32(f32_up(a)), fb = wmax_nonneg_f32(f32_up(b)), fc = wmax_nonneg_f32(f32_up(c));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(WFULL, s, o);
    a = (double)fa;
    b = (double)fb;
    c = (double)fc;
}
// same with the lane-local maxima already in fp32 (rounded up value by value: rounding is monotone, so the result is
// the one above bit for bit, without fp64 maximum sequences on the way)
__device__ __forceinline__ void wred_mmms_f(float fa, float fb, float fc, double& a, double& b, double& c, double& s) {
    fa = wmax_nonneg_f32(fa);
    fb = wmax_nonneg_f32(fb);
    fc = wmax_nonneg_f32(fc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(WFULL, s, o);
    a = (double)fa;
    b = (double)fb;
    c = (double)fc;
}

// one maximum and one sum in one butterfly; the maximum (rho = 1 / step to the boundary) in fp32 rounded UP, which
// can only shorten the step (by < 1.2e-7 relative, inside the 1e-6 margin of the fraction to the boundary)
__device__ __forceinline__ void wred_ms(double& a, double& s) {
    const float fa = wmax_nonneg_f32(f32_up(a));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(WFULL, s, o);
    a = (double)fa;
}

#ifndef BMPC_WARP_MINB
// resident warps (one-warp CTAs) per SM the register allocation is capped for.  Measured on C1 (tools/studies/nscale.py):
// 16 warps x 128 registers 0.181 ms, 12 x 168 registers 0.172 ms, 8 x 228 registers 0.200 ms -- the iteration is a
// dependent chain whose latency grows with the number of co-resident warps (shared-memory / shuffle pipe), so registers
// (batched loads, nothing rematerialised) buy more than residency
#define BMPC_WARP_MINB 12
#endif
template <int NT, int RPL>
__global__ void __launch_bounds__(32, BMPC_WARP_MINB)
    step_warp(const __grid_constant__ StepParams P, const __grid_constant__ WarpParams Q) {
    using D = WarpDims<NT>;
    constexpr int NB = D::NB, LDG = D::LDG, LDH = D::LDH, LDN = D::LDN, LDP = D::LDP;
    constexpr int MP = 32 * RPL, NV2 = LDH / 2;
    extern __shared__ __align__(128) double smem[];
    const WarpLayout& L = Q.L;
    const int lane = threadIdx.x;
    const int fg = lane >> 2, ft = lane & 3;  // DMMA fragment coordinates
    const int half = lane >> 4, i16 = lane & 15;
    const int nz = P.nz, nr = P.n;  // real sizes: move variables, move variables + slack
    const int nY = P.nY, nu = P.nu, ny = P.ny, nx = P.nx, nd = P.nd;
    const int nS = P.rt.nS, nDr = P.rt.nDr, m = Q.m, mD = Q.mD;
    using SM = WarpSmem<NT, RPL>;
    double* sG = smem + SM::G;
    double* sH = smem + SM::H;
    double* sL = smem + SM::L;
    double* sPhi = smem + SM::PHI;  // C tiles of Phi (+ the rhs row), then the factor's columns
    double* vx = smem + SM::VX;
    double* vy = smem + SM::VY;
    double* w1 = smem + SM::W1;
    double* w2 = smem + SM::W2;
    double* wd = smem + SM::WD;
    double* sF = smem + L.F;
    double* stY = smem + L.tY;
    double* sfx = smem + L.fx;
    double* sxh = smem + L.xh;
    double* slu = smem + L.lu;
    double* sd0 = smem + L.dd;
    double* sDh = smem + L.Dh;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);

    if (lane == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (lane < 16) {
        vx[lane] = 0.0;
        vy[lane] = 0.0;
    }
    if (lane < 2) {  // the "no unit row" position MP reads as zero in every row-weight array
        w1[MP + lane] = 0.0;
        w2[MP + lane] = 0.0;
        wd[MP + lane] = 0.0;
    }
    __syncwarp();
    uint32_t phase = 0;
    const bool isvar = lane < NT;   // lane owns a (real or dummy) variable
    const bool isreal = lane < nz;  // ... a real move variable
    const bool iseps = P.neps && lane == nz;
    const int iv = isvar ? lane : 0;
    const double2* vx2 = reinterpret_cast<const double2*>(vx);
    const double2* hrow = reinterpret_cast<const double2*>(sH + iv * LDH);
    const int mh = ((mD + 1) / 2 + 3) & ~3;  // dense rows per half-warp in the G'w products (multiple of 4)
    const int KS = (mD + 3) / 4;              // DMMA k-steps
    const double minv = 1.0 / (double)max(m, 1);
    // per-CTA constants: this lane's positions and, for a variable lane, its unit rows
    int pflag[RPL], prow[RPL], pz3[RPL], pvar[RPL];
    double psig[RPL];
    bool okR[RPL];
#pragma unroll
    for (int t = 0; t < RPL; ++t) {
        const int4 pi = Q.pinfo[lane + 32 * t];
        prow[t] = pi.x;
        pflag[t] = pi.y;
        pz3[t] = pi.z;
        pvar[t] = pi.w;
        okR[t] = (pi.y & PI_VALID) != 0;
        psig[t] = (pi.y & PI_NEG) ? -1.0 : 1.0;
    }
    const int up0 = isvar ? Q.upos[lane] : MP, up1 = isvar ? Q.upos[16 + lane] : MP;

    gather_wait_acks(P, lane);
    PCLK_DECL;
    bool first_pass = true;
    for (;;) {
        PCLK(23);
        // work queue: the first instance of every CTA is its block index (no 2368-way atomic storm at launch), the
        // rest come from the shared counter.  (A pull kernel of an earlier period can only be resident here while it
        // waits for a slower peer, i.e. when this rank has slack anyway, so the static first slot stays on with the
        // fused gather: measured 2-3 us per period better than all-dynamic at 2 GPUs.)
        const bool static_first = Q.static_first != 0;
        int slot = 0;
        if (first_pass && static_first) {
            slot = (int)blockIdx.x;
        } else {
            if (lane == 0) slot = (static_first ? (int)gridDim.x : 0) + (int)atomicAdd(&P.counters[0], 1u);
            slot = __shfl_sync(WFULL, slot, 0);
        }
        first_pass = false;
        if (__all_sync(WFULL, slot >= P.N)) break;  // vote: provably warp-uniform branch (no divergent-shuffle paths)
        const int inst = Q.order ? Q.order[slot] : slot;
        // the instance some CTA will fetch about one instance-time from now (one wave of the queue ahead): its constants are
        // pulled into L2 below, so that CTA's prologue and TMA copy find them there instead of in HBM
        const int pslot = slot + (int)gridDim.x;
        const int pinst = (Q.l2_prefetch && pslot < P.N) ? (Q.order ? Q.order[pslot] : pslot) : -1;

#ifdef BMPC_PHASE_CLK
        if (inst < 0) break;  // (consumes `inst`: the clock below then reads after the order entry has arrived)
#endif
        PCLK(21);
        // ---- stage 0: TMA bulk loads of this instance's matrices ----
        const int lv_ok = P.lv_ok[P.sH ? inst : 0];
        if (lane == 0) {
            fence_proxy_async();
            const uint32_t bG = (uint32_t)(Q.GR * LDG) * 8u, bH = (uint32_t)(2 * NT * LDH) * 8u;
            mbar_arrive_expect_tx(bar, bG + bH);
            tma_bulk_g2s(sH, Q.HL + (long)inst * Q.sHL, bH, bar);
            if (bG) tma_bulk_g2s(sG, Q.Gw + (long)inst * Q.sGw, bG, bar);
        }
        // ---- stage 1: initpred!  (execute.jl:247-277) ----
        // Every global read of the prologue is ISSUED before the first one is consumed (one DRAM round trip instead of a
        // dozen dependent ones): the state, the previous solution and multipliers (warm start), this lane's bounds, and
        // the first 8 / 4 columns of K / V for the lane's two prediction rows.
        const double* gxh = P.est_on ? P.xstate : P.xhat0;
        const double* gK = P.K + (long)inst * P.sK;
        const double* gV = P.V + (long)inst * P.sV;
        const double* gB = P.B + (long)inst * P.sB;
        const double* gyop = P.yop + (long)inst * P.syop;
        const double* guop = P.uop + (long)inst * P.suop;
        const double* gM = P.Mw + (long)inst * P.sM;
        const double* gdb = P.dbound + (long)inst * nDr;
        const double* gsb = P.sbase + (long)inst * nS;
        const double* glw0 = P.lam_ws + (long)inst * P.ws_stride;
        const double pxh = lane < nx ? gxh[(long)inst * nx + lane] : 0.0;
        const double plu = lane < nu ? P.lastu0[(long)inst * nu + lane] : 0.0;
        const double pz = lane < nr ? P.Z[(long)inst * nr + lane] : 0.0;  // previous period's Z̃ (warm start, shifted solution)
        const int ws_on = P.use_ws ? P.ws_flag[inst] : 0;
        double plam[RPL], pbnd[RPL];
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            plam[t] = (P.use_ws && okR[t]) ? glw0[lane + 32 * t] : 0.0;  // (multipliers are kept by position)
            pbnd[t] = okR[t] ? ((pflag[t] & PI_SPARSE) ? gsb[prow[t]] : gdb[prow[t] - nS]) : 0.0;
        }
        constexpr int KC = 8, VC = 4;
        double kpre[2][KC], vpre[2][VC], bpre[2], mpre[2], rypre[2], yoppre[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = lane + 32 * h;
            const bool okt = t < nY;
            const int tc = okt ? t : 0;
#pragma unroll
            for (int k = 0; k < KC; ++k) kpre[h][k] = (okt && k < nx) ? gK[tc + (long)nY * k] : 0.0;
#pragma unroll
            for (int k = 0; k < VC; ++k) vpre[h][k] = (okt && k < nu) ? gV[tc + (long)nY * k] : 0.0;
            bpre[h] = okt ? gB[tc] + (P.Ys ? P.Ys[(long)inst * nY + tc] : 0.0) : 0.0;
            mpre[h] = okt ? gM[tc] : 0.0;
            rypre[h] = okt ? (P.Rhat_y ? P.Rhat_y[(long)inst * nY + tc] : P.ry[(long)inst * ny + (tc % ny)]) : 0.0;
            yoppre[h] = okt ? gyop[tc % ny] : 0.0;
        }
        PCLK(22);
        if (lane < nx) sxh[lane] = pxh;
        for (int k = lane + 32; k < nx; k += 32) sxh[k] = gxh[(long)inst * nx + k];
        if (lane < nu) slu[lane] = plu;
        if (nd > 0) {
            for (int k = lane; k < nd; k += 32) sd0[k] = P.d0[(long)inst * nd + k];
            for (int k = lane; k < nd * P.Hp; k += 32)
                sDh[k] = P.Dhat0 ? P.Dhat0[(long)inst * nd * P.Hp + k] : P.d0[(long)inst * nd + (k % nd)];
        }
        __syncwarp();
        PCLK(18);
        if (__any_sync(WFULL, P.est_on != 0)) skf_correct(P, inst, lane, 32, sxh, sd0, smem + L.ev, [] { __syncwarp(); });
        double racc = 0.0;
        // measured-disturbance terms and the bookkeeping every prediction row shares
        auto finish_row = [&](int t, double f, double mt, double ryt, double yopt) {
            if (nd > 0) {
                const double* gG = P.G + (long)inst * P.sG;
                const double* gJ = P.J + (long)inst * P.sJ;
#pragma unroll 1
                for (int k = 0; k < nd; ++k) f = fma(gG[t + (long)nY * k], sd0[k], f);
#pragma unroll 1
                for (int k = 0; k < nd * P.Hp; ++k) f = fma(gJ[t + (long)nY * k], sDh[k], f);
            }
            sF[t] = f;
            const double cy = f + yopt - ryt;
            const double ty = mt * cy;
            stY[t] = ty;
            racc = fma(cy, ty, racc);
            P.F_out[(long)inst * nY + t] = f;
        };
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = lane + 32 * h;
            if (t < nY) {
                double f = bpre[h];
#pragma unroll
                for (int k = 0; k < KC; ++k) f = fma(kpre[h][k], sxh[k < nx ? k : 0], f);
#pragma unroll 1
                for (int k = KC; k < nx; ++k) f = fma(gK[t + (long)nY * k], sxh[k], f);
#pragma unroll
                for (int k = 0; k < VC; ++k) f = fma(vpre[h][k], slu[k < nu ? k : 0], f);
#pragma unroll 1
                for (int k = VC; k < nu; ++k) f = fma(gV[t + (long)nY * k], slu[k], f);
                finish_row(t, f, mpre[h], rypre[h], yoppre[h]);
            }
        }
#pragma unroll 1
        for (int t = lane + 64; t < nY; t += 32) {
            double f = gB[t] + (P.Ys ? P.Ys[(long)inst * nY + t] : 0.0);
#pragma unroll 2
            for (int k = 0; k < nx; ++k) f = fma(gK[t + (long)nY * k], sxh[k], f);
#pragma unroll 1
            for (int k = 0; k < nu; ++k) f = fma(gV[t + (long)nY * k], slu[k], f);
            finish_row(t, f, gM[t], P.Rhat_y ? P.Rhat_y[(long)inst * nY + t] : P.ry[(long)inst * ny + (t % ny)], gyop[t % ny]);
        }
        PCLK(19);
        if (pinst >= 0) {  // (placed after the first use of this instance's own loads: pinst has arrived by now)
            auto pf = [&](const double* base, long stride, int count) {
                if (stride == 0 || count <= 0) return;  // shared by all instances: resident anyway
                const size_t lo = (size_t)(base + (long)pinst * stride), hi = lo + (size_t)count * 8;
                for (size_t a = (lo & ~(size_t)127) + (size_t)lane * 128; a < hi; a += 32 * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            };
            pf(P.K, P.sK, nY * nx);
            pf(P.V, P.sV, nY * nu);
            pf(P.B, P.sB, nY);
            pf(P.Mw, P.sM, nY);
            pf(P.dbound, nDr, nDr);
            pf(P.sbase, nS, nS);
            if (P.use_ws) pf(P.lam_ws, P.ws_stride, m);
            if (lane == 0) {
                const uint32_t bG = (uint32_t)(Q.GR * LDG) * 8u, bH = (uint32_t)(2 * NT * LDH) * 8u;
                if (Q.sHL) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(Q.HL + (long)pinst * Q.sHL), "r"(bH) : "memory");
                if (Q.sGw && bG) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(Q.Gw + (long)pinst * Q.sGw), "r"(bG) : "memory");
            }
        }
        if (P.has_terminal) {
            const double* gkx = P.kx + (long)inst * P.skx;
            const double* gvx = P.vx + (long)inst * P.svx;
            const double* gbx = P.bx + (long)inst * P.sbx;
#pragma unroll 1
            for (int i = lane; i < nx; i += 32) {
                double f = gbx[i];
#pragma unroll 1
                for (int k = 0; k < nx; ++k) f = fma(gkx[i + (long)nx * k], sxh[k], f);
#pragma unroll 1
                for (int k = 0; k < nu; ++k) f = fma(gvx[i + (long)nx * k], slu[k], f);
                if (nd > 0) {
                    const double* ggx = P.gx + (long)inst * P.sgx;
                    const double* gjx = P.jx + (long)inst * P.sjx;
#pragma unroll 1
                    for (int k = 0; k < nd; ++k) f = fma(ggx[i + (long)nx * k], sd0[k], f);
#pragma unroll 1
                    for (int k = 0; k < nd * P.Hp; ++k) f = fma(gjx[i + (long)nx * k], sDh[k], f);
                }
                sfx[i] = f;
            }
        }
        __syncwarp();
        if (P.nw > 0) {  // linconstraint_custom! (execute.jl:337-366)
            custom_fw(P, inst, lane, 32, sxh, slu, sd0, sDh, sF, smem + L.Fw);
            __syncwarp();
        }
        PCLK(20);
        // ---- linconstraint!  (transcription.jl:811-848): right-hand sides of this lane's rows ----
        double hR[RPL], sR[RPL], lamR[RPL];
        double hmax = 0.0;
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            double hv = 0.0, wq = 0.0;
            if (okR[t]) {
                if (pflag[t] & PI_SPARSE) {
                    const int ch = pz3[t];
                    hv = pbnd[t] - (ch >= 0 ? psig[t] * slu[ch] : 0.0);
                } else {
                    const int src = pz3[t];
                    const double fsrc = src < nY ? sF[src] : (src < nY + nx ? sfx[src - nY] : (smem + L.Fw)[src - nY - nx]);
                    hv = psig[t] * (pbnd[t] - fsrc);
                    // q = 2 Ev' tY = Gt' w with w = 2 sigma tY on ONE row per prediction (the max-side row if present)
                    if (pflag[t] & PI_ISQ) wq = 2.0 * psig[t] * stY[src];
                }
            }
            hR[t] = hv;
            hmax = fmax(hmax, fabs(hv));
            w1[lane + 32 * t] = wq;
        }
        PCLK(12);
        mbar_wait(bar, phase);
        phase ^= 1u;
        PCLK(13);

        // helpers -------------------------------------------------------------------------
        // out[t] = (row at this lane's position t) . v   with v in vx: dense rows from Gt, unit rows +-v_i
        auto row_products = [&](double (&out)[RPL]) {
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                const int p = lane + 32 * t;
                double r = 0.0;
                if (p < mD) {
                    const double2* row = reinterpret_cast<const double2*>(sG + p * LDG);
                    double2 g2[NV2], v2[NV2];
#pragma unroll
                    for (int jj = 0; jj < NV2; ++jj) {  // (all loads first: one shared-memory latency per row, not twelve)
                        g2[jj] = row[jj];
                        v2[jj] = vx2[jj];
                    }
                    double a = 0.0, b = 0.0;
#pragma unroll
                    for (int jj = 0; jj < NV2; ++jj) {
                        a = fma(g2[jj].x, v2[jj].x, a);
                        b = fma(g2[jj].y, v2[jj].y, b);
                    }
                    r = a + b;
                } else if (pflag[t] & PI_UNIT) {
                    r = psig[t] * vx[pvar[t]];
                }
                out[t] = r;
            }
        };
        // (Gt' w)_i for lane i < NT; w in shared memory by position (zero on padding positions)
        const double* gcol = sG + (long)half * mh * LDG + (i16 < NT ? i16 : 0);
        auto gt_apply1 = [&](const double* w) -> double {
            const double2* wv = reinterpret_cast<const double2*>(w + half * mh);
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 2
            for (int r = 0; r < mh; r += 4) {
                const double2 wa = wv[r >> 1], wb = wv[(r >> 1) + 1];
                const double g0 = gcol[r * LDG], g1 = gcol[(r + 1) * LDG], g2 = gcol[(r + 2) * LDG], g3 = gcol[(r + 3) * LDG];
                a0 = fma(g0, wa.x, a0);
                a1 = fma(g1, wa.y, a1);
                a2 = fma(g2, wb.x, a2);
                a3 = fma(g3, wb.y, a3);
            }
            double a = (a0 + a1) + (a2 + a3);
            a += __shfl_xor_sync(WFULL, a, 16);
            return a + (w[up0] - w[up1]);  // unit rows of this lane's variable (position MP reads as zero)
        };
        auto gt_apply2 = [&](const double* wa_, const double* wb_, double& outa, double& outb) {
            const double2* wva = reinterpret_cast<const double2*>(wa_ + half * mh);
            const double2* wvb = reinterpret_cast<const double2*>(wb_ + half * mh);
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll 2
            for (int r = 0; r < mh; r += 4) {  // mh is a multiple of 4
                const double2 wa = wva[r >> 1], wa2 = wva[(r >> 1) + 1], wb = wvb[r >> 1], wb2 = wvb[(r >> 1) + 1];
                const double g0 = gcol[r * LDG], g1 = gcol[(r + 1) * LDG], g2 = gcol[(r + 2) * LDG], g3 = gcol[(r + 3) * LDG];
                a0 = fma(g0, wa.x, a0);
                a1 = fma(g1, wa.y, a1);
                a2 = fma(g2, wa2.x, a2);
                a3 = fma(g3, wa2.y, a3);
                b0 = fma(g0, wb.x, b0);
                b1 = fma(g1, wb.y, b1);
                b2 = fma(g2, wb2.x, b2);
                b3 = fma(g3, wb2.y, b3);
            }
            double a = (a0 + a1) + (a2 + a3), b = (b0 + b1) + (b2 + b3);
            a += __shfl_xor_sync(WFULL, a, 16);
            b += __shfl_xor_sync(WFULL, b, 16);
            outa = a + (wa_[up0] - wa_[up1]);
            outb = b + (wb_[up0] - wb_[up1]);
        };
        // (H x)_i with x in vx
        auto hess_apply = [&]() -> double {
            double2 h2[NV2], v2[NV2];
#pragma unroll
            for (int jj = 0; jj < NV2; ++jj) {
                h2[jj] = hrow[jj];
                v2[jj] = vx2[jj];
            }
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int jj = 0; jj < NV2; ++jj) {
                a = fma(h2[jj].x, v2[jj].x, a);
                b = fma(h2[jj].y, v2[jj].y, b);
            }
            return isvar ? a + b : 0.0;
        };
        auto bcast = [&](double v, int j) -> double { return __shfl_sync(WFULL, v, j); };

        // q_i = 2 sum_t Ev[t,i] tY[t] (+ input-setpoint term)
        double q = 0.0;
        {
            const int ir = i16 < nz ? i16 : 0;
            double a;
            if (P.pd_is_ev) {
                // the dense rows of Gt ARE +-Ev: no second pass over Ev in HBM
                __syncwarp();
                a = 0.5 * gt_apply1(w1);
            } else {
                // column i of Ev from HBM, t range split by half-warp
                const double* col = P.Ev + (long)inst * P.sEv + (long)nY * ir;
                const int th = (nY + 1) >> 1;
                const int t0 = half ? th : 0, t1 = half ? nY : th;
                double a0 = 0.0, a1 = 0.0;
                int t = t0;
#pragma unroll 2
                for (; t + 1 < t1; t += 2) {
                    a0 = fma(col[t], stY[t], a0);
                    a1 = fma(col[t + 1], stY[t + 1], a1);
                }
                if (t < t1) a0 = fma(col[t], stY[t], a0);
                a = a0 + a1;
                a += __shfl_xor_sync(WFULL, a, 16);
            }
            if (P.has_L) {
                const double* gL = P.Lw + (long)inst * P.sL;
                const int l = ir / nu, ch = ir % nu;
#pragma unroll 1
                for (int tt = P.blk_start[l]; tt < P.blk_start[l + 1]; ++tt) {
                    const int idx = tt * nu + ch;
                    const double ru = P.Rhat_u ? P.Rhat_u[(long)inst * P.nU + idx] : guop[ch];
                    a = fma(gL[idx], slu[ch] + guop[ch] - ru, a);
                }
#pragma unroll 1
                for (int idx = lane; idx < P.nU; idx += 32) {
                    const int c2 = idx % nu;
                    const double ru = P.Rhat_u ? P.Rhat_u[(long)inst * P.nU + idx] : guop[c2];
                    const double cu = slu[c2] + guop[c2] - ru;
                    racc = fma(gL[idx] * cu, cu, racc);
                }
            }
            q = isreal ? 2.0 * a : 0.0;
        }
        const double rconst = wsum(racc);
        double qabs = fabs(q);
        {
            double z0 = 0.0, z1 = 0.0;
            wred_mmms(hmax, qabs, z0, z1);
        }
        const double hscale = 1.0 + hmax;
        const double qs = 1.0 + qabs;
        // ---- stage 2: unconstrained minimiser with the cached factor (explicitmpc.jl:209) ----
        double x = 0.0;
        if (__any_sync(WFULL, lv_ok != 0)) {
            double b = isvar ? -q : 0.0;
            const double invd0 = sL[iv * LDH + iv];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double yj = bcast(b * invd0, j);
                if (lane > j && isvar) b = fma(-sL[iv * LDH + j], yj, b);
            }
            b *= invd0;
#pragma unroll
            for (int j = NT - 1; j >= 0; --j) {
                const double xj = bcast(b * invd0, j);
                if (lane < j) b = fma(-sL[j * LDH + iv], xj, b);
            }
            x = isvar ? b * invd0 : 0.0;
        }
        __syncwarp();
        if (lane < 16) vx[lane] = x;
        __syncwarp();
        double gx[RPL];
        row_products(gx);
        double smin = 1e300;
#pragma unroll
        for (int t = 0; t < RPL; ++t) {
            sR[t] = hR[t] - gx[t];
            if (okR[t]) smin = fmin(smin, sR[t]);
        }
        smin = (m > 0) ? wmin(smin) : 0.0;
        const bool feasible = (m == 0) || (lv_ok && smin >= -1e-12 * hscale);
        PCLK(14);
        int status = ST_OPTIMAL, iters = 0;
        double kk0 = 0.0, kk1 = 0.0, kk2 = 0.0;  // relative KKT residuals of the returned iterate (io.kkt)
        if (__any_sync(WFULL, !feasible)) {
            // ---- stage 3: Mehrotra predictor-corrector ----
            const double mu0 = fmax(1e-2 * qs * hscale * minv, 1e-8);
#pragma unroll
            for (int t = 0; t < RPL; ++t) {
                sR[t] = fmax(sR[t], 1e-2 * hscale);
                lamR[t] = okR[t] ? mu0 / sR[t] : 0.0;
            }
            // ---- warm start (set_warmstart_mpc! analogue): previous period's solution and multipliers ----
            if (__any_sync(WFULL, ws_on != 0)) {
                // level coordinates of the previous Z̃: v_l = sum of the moves up to block l of the same input (pz holds Z̃_lane)
                double xw = 0.0;
                for (int k = 0; k * nu < NT; ++k) {  // (uniform trip count)
                    const int src = lane - k * nu;
                    const double zl = __shfl_sync(WFULL, pz, src >= 0 ? src : 0);
                    if (isreal && src >= 0) xw += zl;
                }
                if (iseps) xw = pz;
                __syncwarp();
                if (lane < 16) vx[lane] = xw;
                __syncwarp();
                x = xw;
                row_products(gx);
                const double lmin = 1e-4 * qs / hscale;
#pragma unroll
                for (int t = 0; t < RPL; ++t) {
                    sR[t] = fmax(hR[t] - gx[t], 1e-2 * hscale);
                    lamR[t] = okR[t] ? fmax(plam[t], lmin) : 0.0;
                }
            }
            status = ST_ITERATION_LIMIT;
            double best_merit = 1e300, rp_inf = 0.0;
            int stall = 0;
            // r_p = Gx + s - h is carried by its recurrence r_p <- (1 - a) r_p (ds = -r_p - G dx is formed from the
            // COMPUTED dx, so the recurrence is exact up to rounding whatever the accuracy of the linear solve)
            double rpR[RPL];
#pragma unroll
            for (int t = 0; t < RPL; ++t) rpR[t] = okR[t] ? gx[t] + sR[t] - hR[t] : 0.0;
            const double inv_tol = 1.0 / P.tol, inv_tol_p = 1.0 / (P.tol * hscale), inv_tol_mu = 1.0 / (P.tol_mu * qs * hscale);
            PCLK(15);
            for (int it = 0; it <= P.max_iter; ++it) {
                PCLK(11);
                double dR[RPL], isR[RPL];
                double e_p = 0.0, musum = 0.0;
                float e_pf = 0.f;
#pragma unroll
                for (int t = 0; t < RPL; ++t) {
                    const int p = lane + 32 * t;
                    isR[t] = okR[t] ? rcp_fast(okR[t] ? sR[t] : 1.0) : 0.0;  // (a padding row's s = 0 has no reciprocal)
                    dR[t] = lamR[t] * isR[t];
                    w1[p] = lamR[t];
                    w2[p] = dR[t] * rpR[t];
                    wd[p] = dR[t];
                    e_pf = fmaxf(e_pf, f32_up(fabs(rpR[t])));
                    musum = fma(sR[t], lamR[t], musum);
                }
                __syncwarp();
                PCLK(0);
                // residuals: rd = Hx + q + G'lam; the predictor's right-hand side needs G'(d rp) -- one pass for both
                const double Hxq = hess_apply() + q;
                double Gtl, Gdr;
                gt_apply2(w1, w2, Gtl, Gdr);
                if (!isvar) {
                    Gtl = 0.0;
                    Gdr = 0.0;
                }
                PCLK(1);
                const double rd = Hxq + Gtl;
                double e_d, dsc;
                wred_mmms_f(f32_up(fabs(rd)), fmaxf(f32_up(fabs(Hxq)), f32_up(fabs(Gtl))), e_pf, e_d, dsc, e_p, musum);
                const double qd = qs + dsc;  // scale of the dual residual's terms
                const double mu = musum * minv;
                const double imu = rcp_fast(fmax(mu, 1e-300));  // (needed after the affine step: off its dependent chain)
                rp_inf = e_p;
                if (__any_sync(WFULL, !(e_d == e_d) || !(e_p == e_p) || !(mu == mu) || e_d > 1e250 || e_p > 1e250)) {
                    status = ST_INFEASIBLE;
                    break;
                }
                // (the scales of the primal residual and of the gap are loop invariants: multiplications, one reciprocal)
                kk1 = e_d * rcp_fast(qd);  // (qd >= 1)
                const double merit = fmax(fmax(kk1 * inv_tol, e_p * inv_tol_p), musum * inv_tol_mu);
                kk0 = e_p * inv_tol_p * P.tol;
                kk2 = musum * inv_tol_mu * P.tol_mu;
                if (__any_sync(WFULL, merit <= 1.0 || (best_merit <= 1e3 && merit >= best_merit) ||
                                              (it == P.max_iter && merit <= 1e3))) {
                    status = ST_OPTIMAL;
                    break;
                }
                best_merit = fmin(best_merit, merit);
                if (it == P.max_iter) break;
                iters = it + 1;
                PCLK(2);
                // ---- Phi = H + Gt' D Gt on the FP64 tensor pipe (lower block-triangle) ----
                double c00a, c00b, c10a = 0.0, c10b = 0.0, c11a = 0.0, c11b = 0.0;
                {
                    const int hr = fg < NT ? fg : 0, hc = 2 * ft;
                    const double2 h2 = *reinterpret_cast<const double2*>(sH + hr * LDH + (hc < LDH ? hc : 0));
                    const bool okr = fg < NT;
                    c00a = (okr && hc < NT) ? h2.x : 0.0;
                    c00b = (okr && hc + 1 < NT) ? h2.y : 0.0;
                    if (NB == 2) {
                        const int hr1 = 8 + fg < NT ? 8 + fg : 0;
                        const bool okr1 = 8 + fg < NT;
                        const double2 g2 = *reinterpret_cast<const double2*>(sH + hr1 * LDH + (hc < LDH ? hc : 0));
                        c10a = (okr1 && hc < NT) ? g2.x : 0.0;
                        c10b = (okr1 && hc + 1 < NT) ? g2.y : 0.0;
                        const int hc1 = 8 + hc;
                        const double2 k2 = *reinterpret_cast<const double2*>(sH + hr1 * LDH + (hc1 < LDH ? hc1 : 0));
                        c11a = (okr1 && hc1 < NT) ? k2.x : 0.0;
                        c11b = (okr1 && hc1 + 1 < NT) ? k2.y : 0.0;
                    }
                }
                {
                    constexpr bool pred1 = NB == 2 && LDG < 16;  // columns 8+fg exist only for fg < LDG - 8
                    const bool ok1 = !pred1 || fg < LDG - 8;
                    const double* ga = sG + ft * LDG + fg;
                    const double* gb = sG + ft * LDG + (ok1 ? 8 + fg : 0);
                    const double* wdp = wd + ft;
                    // software pipeline: the fragments of k-step ks+1 are loaded before the DMMAs of k-step ks issue
                    double a0 = 0.0, a1 = 0.0, dk = 0.0;
                    if (KS > 0) {
                        a0 = ga[0];
                        dk = wdp[0];
                        if (NB == 2) a1 = gb[0];
                    }
#pragma unroll 2
                    for (int ks = 0; ks < KS; ++ks) {
                        const int kn = ks + 1 < KS ? ks + 1 : ks;
                        const double a0n = ga[kn * 4 * LDG];
                        const double dkn = wdp[kn * 4];
                        double a1n = 0.0;
                        if (NB == 2) a1n = gb[kn * 4 * LDG];
                        const double b0 = dk * a0;
                        dmma884(c00a, c00b, a0, b0);
                        if (NB == 2) {
                            if (pred1) a1 = ok1 ? a1 : 0.0;
                            const double b1 = dk * a1;
                            dmma884(c10a, c10b, a1, b0);
                            dmma884(c11a, c11b, a1, b1);
                        }
                        a0 = a0n;
                        a1 = a1n;
                        dk = dkn;
                    }
                }
                __syncwarp();
                *reinterpret_cast<double2*>(sPhi + fg * LDP + 2 * ft) = make_double2(c00a, c00b);
                if (NB == 2) {
                    *reinterpret_cast<double2*>(sPhi + (8 + fg) * LDP + 2 * ft) = make_double2(c10a, c10b);
                    *reinterpret_cast<double2*>(sPhi + (8 + fg) * LDP + 8 + 2 * ft) = make_double2(c11a, c11b);
                }
                // the predictor's right-hand side rides along as an extra row of the matrix being factored
                const double rhs0 = isvar ? -(Hxq + Gdr) : 0.0;
                if (lane < LDP) sPhi[8 * NB * LDP + lane] = rhs0;  // (row 8 NB: past the C tiles)
                const double dunit = wd[up0] + wd[up1];  // unit rows: a diagonal term of Phi
                __syncwarp();
                // rows of the bordered matrix: lane < NT row `lane` of Phi, lane NT the predictor's right-hand side
                double phi[2 * NV2];
                {
                    const double2* prw = reinterpret_cast<const double2*>(sPhi + (lane < NT ? lane : 8 * NB) * LDP);
#pragma unroll
                    for (int jj = 0; jj < NV2; ++jj) {
                        const double2 p2 = prw[jj];
                        phi[2 * jj] = p2.x;
                        phi[2 * jj + 1] = p2.y;
                    }
                }
                // this lane's pivot; the diagonal entry of its register row is never read (the column updates only use the
                // entries below the diagonal), so the unit rows' term goes into the pivot alone
                double pdiag = sPhi[(lane < NT ? lane : 0) * (LDP + 1)] + dunit;
                __syncwarp();
                PCLK(3);
                // ---- Cholesky: right-looking, rows in registers; column k goes through shared memory (one store, then
                // broadcast LDS.128 reads of the entries below the diagonal).  The COLUMN-SCALED factor M = L diag(L)^-1
                // (unit diagonal) is kept in shared memory, column k contiguous, for the substitutions: they then run on
                // raw broadcasts, without a multiply by 1/L_jj on their dependent chains, and nothing of the factor stays
                // in registers during the row phases ----
                // (LDL' form: column k travels through shared memory UNSCALED, so its store, the warp barrier and the
                // broadcast loads are issued while the pivot's reciprocal is still in flight; M = L D^-1 entries above and
                // on the diagonal are stored as zeros, so the substitutions below need no selects)
                double* colbuf = sPhi;
                double* mbuf = sPhi + NT * LDN;
                double* dinv = smem + SM::DV;
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    const double uk = phi[k];
                    if (k + 1 < NT) {
                        if (lane < NT) colbuf[k * LDN + lane] = uk;
                        __syncwarp();
                    }
                    const double dk = bcast(pdiag, k);
                    // guarded pivot (non-positive or NaN -> huge pivot = the variable is frozen)
                    const double rk = (dk > 1e-280) ? rcp_fast(dk > 1e-280 ? dk : 1.0) : 1e-200;
                    // column k of M.  Lanes <= k (and the idle lanes) compute finite-or-NaN garbage here: it only ever
                    // lands in their own upper-triangle entries and spent pivots, which nothing reads.
                    const double mik = uk * rk;
                    dinv[k] = rk;  // (same value from every lane)
                    pdiag = fma(-mik, uk, pdiag);
                    if (lane < NT) mbuf[k * LDN + lane] = lane > k ? mik : 0.0;
                    if (k + 1 < NT) {
#pragma unroll
                        for (int jp = (k + 1) / 2; jp <= (NT - 1) / 2; ++jp) {
                            const double2 c2 = *reinterpret_cast<const double2*>(colbuf + k * LDN + 2 * jp);
                            if (2 * jp > k) phi[2 * jp] = fma(-mik, c2.x, phi[2 * jp]);
                            if (2 * jp + 1 < NT) phi[2 * jp + 1] = fma(-mik, c2.y, phi[2 * jp + 1]);
                        }
                    }
                    phi[k] = mik;  // (the rhs lane ends up with z = D^-1 M^-1 rhs: its forward substitution came for free)
                }
                PCLK(4);
                if (lane == NT) {
#pragma unroll
                    for (int jj = 0; jj < NV2; ++jj) *reinterpret_cast<double2*>(vy + 2 * jj) = make_double2(phi[2 * jj], 2 * jj + 1 < NT ? phi[2 * jj + 1] : 0.0);
                }
                __syncwarp();
                const double z0 = isvar ? vy[iv] : 0.0;
                const double invd2 = dinv[iv];  // 1 / d_lane
                PCLK(5);
                // Phi = M D M':  x = M'^-1 D^-1 M^-1 b.  Row `lane` / column `lane` of M are re-read from shared memory at every
                // sweep (independent loads, issued ahead of the dependent shuffle + FMA chain).  A sweep advances TWO variables
                // per step: both raw entries of a column pair are broadcast at once and every lane finishes the second one
                // itself with the pair's coupling entry M[2b+1][2b] -- NT/2 dependent steps of (broadcasts, 2 FMAs), the same
                // operations in the same order as a one-variable-per-step sweep
                auto solve_fwd = [&](double b) -> double {
                    double mr[NT];
#pragma unroll
                    for (int j = 0; j < NT; ++j) mr[j] = mbuf[j * LDN + iv];  // (zero for j >= lane and on the idle lanes: row 0)
#pragma unroll
                    for (int k = 0; k + 1 < NT; k += 2) {  // M^-1 (unit lower); the last column of an odd NT has no entries
                        const double y1 = bcast(b, k), b2 = bcast(b, k + 1);
                        const double y2 = fma(-mbuf[k * LDN + k + 1], y1, b2);
                        b = fma(-mr[k + 1], y2, fma(-mr[k], y1, b));
                    }
                    return b * invd2;
                };
                auto solve_back = [&](double b) -> double {
                    double mc[2 * NV2];
                    const double2* ccol = reinterpret_cast<const double2*>(mbuf + iv * LDN);
#pragma unroll
                    for (int jp = 0; jp < NV2; ++jp) {
                        const double2 c2 = ccol[jp];
                        mc[2 * jp] = c2.x;  // (zero for rows <= lane; the idle lanes read column 0 and are masked below)
                        mc[2 * jp + 1] = c2.y;
                    }
                    if (NT & 1) b = fma(-mc[NT - 1], bcast(b, NT - 1), b);  // (the last column of an odd NT is on its own)
#pragma unroll
                    for (int k = (NT & ~1) - 2; k >= 0; k -= 2) {  // M'^-1 (unit upper)
                        const double x2 = bcast(b, k + 1), b1 = bcast(b, k);
                        const double x1 = fma(-mbuf[k * LDN + k + 1], x2, b1);
                        b = fma(-mc[k], x1, fma(-mc[k + 1], x2, b));
                    }
                    return isvar ? b : 0.0;
                };
                // ---- predictor (pass 0) and corrector (pass 1) share one copy of the code ----
                double dsR[RPL], dlR[RPL], rcR[RPL];
#pragma unroll
                for (int t = 0; t < RPL; ++t) rcR[t] = 0.0;
                double rhs = 0.0;
                double dx = 0.0, rho = 0.0, ratio = 1.0;
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    double zz = z0;
                    if (pass) zz = solve_fwd(rhs);
                    dx = solve_back(zz);
                    PCLK(6 + 3 * pass);
                    __syncwarp();
                    if (lane < 16) vx[lane] = dx;
                    __syncwarp();
                    row_products(gx);
                    float rhof = 0.f;  // max over rows of -ds/s and -dl/lambda (step to the boundary = 1/rho), rounded up
                    double sdd = 0.0;  // sum ds*dl
#pragma unroll
                    for (int t = 0; t < RPL; ++t) {
                        dsR[t] = -rpR[t] - gx[t];
                        const double rs_ = -dsR[t] * isR[t];
                        double rl;
                        if (pass == 0) {
                            dlR[t] = -lamR[t] - dR[t] * dsR[t];
                            rl = okR[t] ? 1.0 - rs_ : 0.0;  // -dl/lam = 1 + ds/s in the affine step
                        } else {
                            dlR[t] = -(rcR[t] + lamR[t] * dsR[t]) * isR[t];
                            // (a padding row's lam = 0 would send the whole warp through drcp's slow path every iteration; the
                            // empty asm keeps the compiler from folding the guard back into the outer select)
                            double lden = okR[t] ? lamR[t] : 1.0;
                            asm volatile("" : "+d"(lden));
                            rl = okR[t] ? -dlR[t] * rcp_fast(lden) : 0.0;
                        }
                        rhof = fmaxf(rhof, fmaxf(f32_up(rs_), f32_up(rl)));
                        sdd = fma(dsR[t], dlR[t], sdd);
                    }
                    rho = (double)wmax_nonneg_f32(rhof);
                    if (pass == 0) sdd = wsum(sdd);  // (the sum is only needed after the affine step)
                    PCLK(7 + 3 * pass);
                    if (pass == 0) {
                        // affine step: s*dl + lam*ds = -s*lam exactly, so
                        //   sum (s + a ds)(lam + a dl) = (1 - a) sum s*lam + a^2 sum ds*dl   -- no second reduction
                        const double a_aff = rho > 1.0 ? rcp_fast(rho) : 1.0;
                        const double mua = fmax(((1.0 - a_aff) * musum + a_aff * a_aff * sdd) * minv, 0.0);
                        ratio = mua * imu;
                        const double sig = ratio * ratio * ratio;
#pragma unroll
                        for (int t = 0; t < RPL; ++t) {
                            rcR[t] = sR[t] * lamR[t] + dsR[t] * dlR[t] - sig * mu;
                            w1[lane + 32 * t] = (lamR[t] * rpR[t] - rcR[t]) * isR[t];
                        }
                        __syncwarp();
                        const double Gw = gt_apply1(w1);
                        rhs = isvar ? -(rd + Gw) : 0.0;
                        PCLK(8);
                    }
                }
                // fraction to the boundary: 0.99, tending to 1 as the affine step closes the gap
                const double tau = fmin(fmax(0.99, 1.0 - ratio), 1.0 - 1e-6);  // never exactly onto the boundary
                const double a = rho > tau ? tau * rcp_fast(rho) : 1.0;
                // infeasibility: two collapsed steps with the primal residual still open end the solve (ipm_solve,
                // bmpc_device.cuh)
                stall = (a < 1e-8 && rp_inf > 1e-6 * hscale) ? stall + 1 : 0;
                if (__any_sync(WFULL, stall >= 2)) {
                    status = ST_INFEASIBLE;
                    break;
                }
                x = fma(a, dx, x);
                const double oma = 1.0 - a;
#pragma unroll
                for (int t = 0; t < RPL; ++t) {
                    rpR[t] *= oma;
                    sR[t] = fma(a, dsR[t], sR[t]);
                    lamR[t] = fma(a, dlR[t], lamR[t]);
                }
                __syncwarp();
                if (lane < 16) vx[lane] = x;
                __syncwarp();
            }
            // (status INFEASIBLE only from the certified exits: collapsed steps or NaN; an iteration-limit exit keeps its iterate)
        }
        PCLK(16);
        // ---- stage 4: getinput!  (execute.jl:536-546) ----
        double* gZ = P.Z + (long)inst * nr;
        __syncwarp();
        if (__any_sync(WFULL, status == ST_INFEASIBLE)) {
            // shifted previous solution (set_warmstart_mpc!, transcription.jl:997-1007), in level coordinates
            double a = 0.0;
            if (isreal)
                for (int l = lane % nu; l <= lane; l += nu) a += (l + nu < nz) ? gZ[l + nu] : 0.0;
            x = isreal ? a : (iseps ? gZ[nz] : 0.0);
        }
        __syncwarp();
        if (lane < 16) vx[lane] = x;
        __syncwarp();
        const double Hx = hess_apply();
        double jacc = isvar ? x * (0.5 * Hx + q) : 0.0;
        jacc = wsum(jacc) + rconst;
        // q̃ in reference coordinates: q̃[j] = sum_{l' >= l(j), same input} q_v[l']
        w1[lane] = q;
        __syncwarp();
        if (P.zg_world > 0 && lane < nr) {  // fused all-gather of Z̃: peer stores over NVLink (slot (rank, inst) of every peer)
            const double zv = isreal ? x - (lane >= nu ? vx[lane - nu] : 0.0) : x;
            const long off = P.zg_base + (long)inst * nr + lane;
            if (P.zg_pull) {
                P.zg[P.zg_rank][off] = zv;
            } else {
                for (int pr = 0; pr < P.zg_world; ++pr) P.zg[pr][off] = zv;
            }
        }
        if (isreal) {
            gZ[lane] = x - (lane >= nu ? vx[lane - nu] : 0.0);
            double a = 0.0;
            for (int l = lane; l < nz; l += nu) a += w1[l];
            P.qt_out[(long)inst * nr + lane] = a;
        }
        if (iseps) {
            gZ[nz] = x;
            P.qt_out[(long)inst * nr + nz] = 0.0;
        }
        if (__any_sync(WFULL, P.est_on != 0)) skf_predict(P, inst, lane, 32, sxh, slu, vx, sd0);
        if (lane < nu) {
            const double du = vx[lane];
            const double lu = slu[lane];
            P.lastu_prev[(long)inst * nu + lane] = lu;
            P.lastu0[(long)inst * nu + lane] = lu + du;
            P.u[(long)inst * nu + lane] = lu + du + guop[lane];
        }
        if (P.use_ws) {
            // multipliers for the next period's warm start (valid only after a converged IPM solve), by position
            const bool keep = status == ST_OPTIMAL && iters > 0;
            double* glw = P.lam_ws + (long)inst * P.ws_stride;
            if (keep) {
#pragma unroll
                for (int t = 0; t < RPL; ++t)
                    if (okR[t]) glw[lane + 32 * t] = lamR[t];
            }
            if (lane == 0) P.ws_flag[inst] = keep ? 1 : 0;
        }
        if (lane == 0) {
            if (P.kkt_out) {
                P.kkt_out[(long)inst * 3] = kk0;
                P.kkt_out[(long)inst * 3 + 1] = kk1;
                P.kkt_out[(long)inst * 3 + 2] = kk2;
            }
            P.r_out[inst] = rconst;
            if (P.J_out) P.J_out[inst] = jacc;
            P.status[inst] = status;
            P.iters[inst] = iters;
            if (Q.order_next) {
                // next launch's order: filed from the back in completion order, so the last to finish start first next time
                // (long_thresh, normally never reached, sends instances with that many iterations to the front instead)
                if (iters >= Q.long_thresh)
                    Q.order_next[atomicAdd(&Q.ocnt[0], 1u)] = inst;
                else
                    Q.order_next[P.N - 1 - (int)atomicAdd(&Q.ocnt[1], 1u)] = inst;
            }
        }
        fence_proxy_async();  // generic-proxy accesses to the TMA destinations precede the next bulk copy
        __syncwarp();
        PCLK(17);
#ifdef BMPC_PHASE_CLK
        if (lane == 0 && Q.clk) { atomicAdd((unsigned long long*)&Q.clk[24], 1ull); atomicAdd((unsigned long long*)&Q.clk[25], (unsigned long long)iters); }
#endif
        PCLK_FLUSH();
    }
    // ---- reset the work counters for the next launch (last CTA out) ----
    __syncwarp();  // the other lanes' (peer) stores are ordered before lane 0's fence
    if (lane == 0) {
        if (P.zg_world > 0 && !P.zg_pull) __threadfence_system(); else __threadfence();
        const unsigned done = atomicAdd(&P.counters[1], 1u);
        if (done == gridDim.x - 1) {
            P.counters[0] = 0u;
            P.counters[1] = 0u;
            if (Q.ocnt) {
                Q.ocnt[0] = 0u;
                Q.ocnt[1] = 0u;
            }
            __threadfence();
            publish_epoch(P);
        }
    }
}

}  // namespace bmpc
