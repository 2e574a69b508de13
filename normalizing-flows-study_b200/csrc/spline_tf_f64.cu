// a5/a6 register-path instantiations, float64
#include "transform_impl.cuh"
namespace nf {
template int spline_transform_launch<double, false, false>(const SplineTfArgs<double>&, cudaStream_t);
template int spline_transform_launch<double, true, false>(const SplineTfArgs<double>&, cudaStream_t);
}
