// scratch: instantiate the paired FFT kernels alone (registers / spills under different launch bounds)
#include "../signalsmith_stretch_b200/csrc/stft2.cuh"
using namespace b200s;
template __global__ void b200s::k_analyse2<3072>(Ctx);
template __global__ void b200s::k_synth2<3072>(Ctx);
