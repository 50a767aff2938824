#include "../signalsmith_stretch_b200/csrc/chain_direct6.cuh"
using namespace b200s;
template __global__ void b200s::k_chain_direct6<4, true, false, 0>(Ctx);
template __global__ void b200s::k_chain_direct6<4, true, false, 1>(Ctx);
template __global__ void b200s::k_chain_direct6<4, true, false, 2>(Ctx);
template __global__ void b200s::k_chain_direct6<4, true, false, 3>(Ctx);
template __global__ void b200s::k_chain_direct6<4, true, false, 4>(Ctx);
