// num_bins in (16,32]: same source, bin loops not unrolled (arrays live in local memory; rarely used)
#define NF_UNROLL _Pragma("unroll 1")
#include "transform_impl.cuh"
namespace nf {
template int rqs_unit_fwd_launch<float, true>(const void*, const void*, const void*, const void*, void*, void*, int64_t, int, int, RqsCfg<float>, cudaStream_t);
template int rqs_unit_bwd_launch<float, true>(const void*, const void*, const void*, const void*, const void*, const void*, void*, void*, void*, void*, int64_t, int, int, RqsCfg<float>, cudaStream_t);
template int rqs_unit_fwd_launch<double, true>(const void*, const void*, const void*, const void*, void*, void*, int64_t, int, int, RqsCfg<double>, cudaStream_t);
template int rqs_unit_bwd_launch<double, true>(const void*, const void*, const void*, const void*, const void*, const void*, void*, void*, void*, void*, int64_t, int, int, RqsCfg<double>, cudaStream_t);
template int spline_transform_launch<float, false, true>(const SplineTfArgs<float>&, cudaStream_t);
template int spline_transform_launch<float, true, true>(const SplineTfArgs<float>&, cudaStream_t);
template int spline_transform_launch<double, false, true>(const SplineTfArgs<double>&, cudaStream_t);
template int spline_transform_launch<double, true, true>(const SplineTfArgs<double>&, cudaStream_t);
}
