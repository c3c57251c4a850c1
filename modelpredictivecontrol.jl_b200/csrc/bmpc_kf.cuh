// Small dense Kalman-filter covariance algebra shared by the MHE arrival-covariance update (bmpc_mhe.cuh) and the
// time-varying KalmanFilter fused into the LinMPC step (bmpc_api.cu): reference src/estimator/kalman.jl:1235-1290.
#pragma once
#include <cuda_runtime.h>

namespace bmpc {

// lower-triangular in-place Cholesky + inverse of a small SPD matrix (n <= 32) by one thread.
__device__ inline bool small_spd_inverse(const double* A, double* inv, double* L, int n) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double a = A[i + n * j];
            for (int p = 0; p < j; ++p) a -= L[i + n * p] * L[j + n * p];
            if (i == j) {
                if (!(a > 0.0) || !isfinite(a)) return false;
                L[i + n * i] = sqrt(a);
            } else {
                L[i + n * j] = a / L[j + n * j];
            }
        }
    // inv = L^-T L^-1, column by column
    for (int c = 0; c < n; ++c) {
        double y[32];
        for (int i = 0; i < n; ++i) {
            double a = (i == c) ? 1.0 : 0.0;
            for (int p = 0; p < i; ++p) a -= L[i + n * p] * y[p];
            y[i] = a / L[i + n * i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double a = y[i];
            for (int p = i + 1; p < n; ++p) a -= L[p + n * i] * inv[p + n * c];
            inv[i + n * c] = a / L[i + n * i];
        }
    }
    return true;
}


// Kalman correction of a covariance (correct_estimate_kf!, kalman.jl:1235-1268): out[0:nx*nx] <- (I - K Cm) P with
// K = P Cm' (Cm P Cm' + R)^-1, raw (not symmetrised).  P (nx x nx, symmetric) in shared memory; scratch: K nx*nym,
// M nym*nym + max(nx,nym)^2, out 2*max(nx,nym)^2 + nx*nym doubles.  Called by every thread of the CTA.
template <class Sync>
__device__ inline void kf_correct_cov(int tid, int nth, Sync sync, int nx, int nym, const double* __restrict__ gCm,
                                      const double* __restrict__ gRm, const double* sP, double* sK, double* sM,
                                      double* out) {
    for (int e = tid; e < nx * nym; e += nth) {  // K <- P Cm'   (nx x nym)
        const int i = e % nx, j = e / nx;
        double a = 0.0;
        for (int k = 0; k < nx; ++k) a = fma(sP[i + nx * k], gCm[j + nym * k], a);
        sK[e] = a;
    }
    sync();
    for (int e = tid; e < nym * nym; e += nth) {  // M <- Cm P Cm' + R
        const int i = e % nym, j = e / nym;
        double a = gRm[e];
        for (int k = 0; k < nx; ++k) a = fma(gCm[i + nym * k], sK[k + nx * j], a);
        sM[e] = a;
    }
    sync();
    if (tid == 0) small_spd_inverse(sM, out, sM + nym * nym, nym);  // out[0:nym^2] = M^-1
    sync();
    for (int e = tid; e < nym * nym; e += nth) sM[e] = out[e];
    sync();
    for (int e = tid; e < nx * nym; e += nth) {  // Kg <- (P Cm') M^-1 into out[nx*nx ...]
        const int i = e % nx, j = e / nx;
        double a = 0.0;
        for (int k = 0; k < nym; ++k) a = fma(sK[i + nx * k], sM[k + nym * j], a);
        out[nx * nx + e] = a;
    }
    sync();
    for (int e = tid; e < nx * nx; e += nth) {  // Pnew = P - Kg (Cm P) ; (Cm P) = (P Cm')' for symmetric P
        const int i = e % nx, j = e / nx;
        double a = sP[e];
        for (int k = 0; k < nym; ++k) {
            double cp = 0.0;  // (Cm P)[k, j]
            for (int l = 0; l < nx; ++l) cp = fma(gCm[k + nym * l], sP[l + nx * j], cp);
            a = fma(-out[nx * nx + i + nx * k], cp, a);
        }
        out[e] = a;
    }
    sync();
}

// Covariance recursion of the time-varying KalmanFilter for NM models (one CTA each), independent of the data:
//   mode 1 (preparestate!, correct_estimate_kf! kalman.jl:1235-1268):  K̂ = P̂ Ĉm' (Ĉm P̂ Ĉm' + R̂)^-1 -> Kout,
//                                                                       P̂ <- Hermitian((I - K̂ Ĉm) P̂, :L)
//   mode 2 (updatestate!,  predict_estimate_kf! :1270-1290):            P̂ <- Hermitian(Â P̂ Â' + Q̂, :L)
// The state update itself (x̂ += K̂ v̂, x̂ <- Â x̂ + ...) runs inside the step kernel with the gain written here.
// Shared memory (doubles): P nx^2 | T1 nx^2 | P2 2 nq^2 + nx nym | K nx nym | M nym^2 + nq^2, nq = max(nx, nym).
static __global__ void k_kf_cov(int NM, int nx, int nym, int mode, const double* __restrict__ A,
                                const double* __restrict__ Qc, const double* __restrict__ Cm,
                                const double* __restrict__ Rm, double* __restrict__ Pall, double* __restrict__ Kout) {
    extern __shared__ double sm[];
    const int inst = blockIdx.x;
    if (inst >= NM) return;
    const int nq = nx > nym ? nx : nym;
    double* P = sm;
    double* T1 = sm + nx * nx;
    double* P2 = T1 + nx * nx;
    double* K = P2 + 2 * nq * nq + nx * nym;
    double* M = K + nx * nym;
    double* gP = Pall + (long)inst * nx * nx;
    for (int e = threadIdx.x; e < nx * nx; e += blockDim.x) P[e] = gP[e];
    __syncthreads();
    if (mode == 1) {
        kf_correct_cov((int)threadIdx.x, (int)blockDim.x, [] { __syncthreads(); }, nx, nym, Cm + (long)inst * nym * nx,
                       Rm + (long)inst * nym * nym, P, K, M, P2);
        for (int e = threadIdx.x; e < nx * nym; e += blockDim.x) Kout[(long)inst * nx * nym + e] = P2[nx * nx + e];
        for (int e = threadIdx.x; e < nx * nx; e += blockDim.x) {
            const int i = e % nx, j = e / nx;
            gP[e] = i >= j ? P2[e] : P2[j + nx * i];
        }
    } else {
        const double* gA = A + (long)inst * nx * nx;
        const double* gQ = Qc + (long)inst * nx * nx;
        for (int e = threadIdx.x; e < nx * nx; e += blockDim.x) {  // T1 = P A'
            const int i = e % nx, j = e / nx;
            double a = 0.0;
            for (int k = 0; k < nx; ++k) a = fma(P[i + nx * k], gA[j + nx * k], a);
            T1[e] = a;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < nx * nx; e += blockDim.x) {  // P2 = A T1 + Q
            const int i = e % nx, j = e / nx;
            double a = gQ[e];
            for (int k = 0; k < nx; ++k) a = fma(gA[i + nx * k], T1[k + nx * j], a);
            P2[e] = a;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < nx * nx; e += blockDim.x) {
            const int i = e % nx, j = e / nx;
            gP[e] = i >= j ? P2[e] : P2[j + nx * i];
        }
    }
}

}  // namespace bmpc
