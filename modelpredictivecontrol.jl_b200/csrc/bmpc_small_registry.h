// Registry of the compiled specialisations of the small-controller kernel (bmpc_small.cuh).
// Each small_inst_XX.cu translation unit instantiates step_small<NZT, NEPS, DS, SS> for one NZT
// (so the units compile in parallel) and registers launchers here.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "bmpc_small.cuh"

namespace bmpc {

typedef cudaError_t (*SmallLaunchFn)(const StepParams&, const SmallParams&, int grid, int smem, cudaStream_t);

struct SmallEntry {
    int nzt, neps, ds, ss;
    SmallLaunchFn launch;
    const void* func;  // kernel symbol, for cudaFuncSetAttribute / occupancy queries
};

template <int NZT, int NEPS, int DS, int SS>
cudaError_t small_launch(const StepParams& P, const SmallParams& Q, int grid, int smem, cudaStream_t s) {
    step_small<NZT, NEPS, DS, SS><<<grid, 64, smem, s>>>(P, Q);
    return cudaGetLastError();
}

template <int NZT, int NEPS, int DS, int SS>
SmallEntry small_entry() {
    return SmallEntry{NZT, NEPS, DS, SS, &small_launch<NZT, NEPS, DS, SS>,
                      reinterpret_cast<const void*>(&step_small<NZT, NEPS, DS, SS>)};
}

void small_register_02(std::vector<SmallEntry>&);
void small_register_04(std::vector<SmallEntry>&);
void small_register_06(std::vector<SmallEntry>&);
void small_register_08(std::vector<SmallEntry>&);
void small_register_10(std::vector<SmallEntry>&);
void small_register_12(std::vector<SmallEntry>&);
void small_register_15(std::vector<SmallEntry>&);

}  // namespace bmpc
