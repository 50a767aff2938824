#include "../signalsmith_stretch_b200/csrc/chain_direct6.cuh"
using namespace b200s;
template __global__ void b200s::k_chain_direct6<4, true, false, 0, true>(Ctx);
template __global__ void b200s::k_chain_direct6<4, true, true, 0, true>(Ctx);
template __global__ void b200s::k_chain_direct6<4, false, false, 0, true>(Ctx);
