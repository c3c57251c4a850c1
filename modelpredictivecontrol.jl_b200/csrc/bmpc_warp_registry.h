// Registry of the compiled specialisations of the warp-per-controller kernel (bmpc_warp.cuh).
// Each warp_inst_XX.cu translation unit instantiates step_warp<NT, RPL> for one NT (so the units
// compile in parallel) and registers launchers here.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "bmpc_warp.cuh"

namespace bmpc {

typedef cudaError_t (*WarpLaunchFn)(const StepParams&, const WarpParams&, int grid, int smem, cudaStream_t);

struct WarpEntry {
    int nt, rpl, ldg, ldh, ldn, ldp;
    int off_g;  // WarpSmem<NT, RPL>::G: where the kernel expects the dense rows (the arrays before it have fixed offsets)
    WarpLaunchFn launch;
    const void* func;  // kernel symbol, for cudaFuncSetAttribute / occupancy queries
};

template <int NT, int RPL>
cudaError_t warp_launch(const StepParams& P, const WarpParams& Q, int grid, int smem, cudaStream_t s) {
    step_warp<NT, RPL><<<grid, 32, smem, s>>>(P, Q);
    return cudaGetLastError();
}

template <int NT, int RPL>
WarpEntry warp_entry() {
    using D = WarpDims<NT>;
    return WarpEntry{NT, RPL, D::LDG, D::LDH, D::LDN, D::LDP, WarpSmem<NT, RPL>::G, &warp_launch<NT, RPL>,
                     reinterpret_cast<const void*>(&step_warp<NT, RPL>)};
}

template <int NT>
void warp_register_nt(std::vector<WarpEntry>& v) {
    v.push_back(warp_entry<NT, 1>());
    v.push_back(warp_entry<NT, 2>());
    v.push_back(warp_entry<NT, 4>());
}

void warp_register_03(std::vector<WarpEntry>&);
void warp_register_05(std::vector<WarpEntry>&);
void warp_register_07(std::vector<WarpEntry>&);
void warp_register_09(std::vector<WarpEntry>&);
void warp_register_11(std::vector<WarpEntry>&);
void warp_register_13(std::vector<WarpEntry>&);
void warp_register_16(std::vector<WarpEntry>&);

}  // namespace bmpc
