// fp16-operand instantiations of the TMA-fed conv kernel (the precision mode, Net.precision = "fp16"): same kernels as
// conv3x3.cu's bf16 ones with the 16-bit format switched; a separate translation unit so the two sets compile in parallel.
#define RRIN_CONV2_INSTANTIATE
#include "conv3x3_launch.cuh"

namespace rrin {

int launch_v2_f16(int cfg, const ConvParamsV2& p, const CUtensorMap& tm0, const CUtensorMap& tm1, const CUtensorMap& tmo,
                  const CUtensorMap& tmw, int grid, cudaStream_t stream) {
    return launch_v2_impl<1>(cfg, p, tm0, tm1, tmo, tmw, grid, stream);
}

}  // namespace rrin
