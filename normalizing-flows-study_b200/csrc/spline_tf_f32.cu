// a5/a6 register-path instantiations, float32
#include "transform_impl.cuh"
namespace nf {
template int spline_transform_launch<float, false, false>(const SplineTfArgs<float>&, cudaStream_t);
template int spline_transform_launch<float, true, false>(const SplineTfArgs<float>&, cudaStream_t);
}
