// scratch: instantiate single chain kernels for a quick look at registers / SASS
#include "../signalsmith_stretch_b200/csrc/chain_direct6.cuh"
using namespace b200s;
template __global__ void b200s::k_chain_direct6<4, true, false>(Ctx);
#ifdef ALSO4
template __global__ void b200s::k_chain_direct4<4, true>(Ctx);
#endif
#ifdef ALSODUAL
template __global__ void b200s::k_chain_direct6<4, true, true>(Ctx);
#endif
