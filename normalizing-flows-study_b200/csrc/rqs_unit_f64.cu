// a7 register-path instantiations, float64
#include "transform_impl.cuh"
namespace nf {
template int rqs_unit_fwd_launch<double, false>(const void*, const void*, const void*, const void*, void*, void*, int64_t, int, int, RqsCfg<double>, cudaStream_t);
template int rqs_unit_bwd_launch<double, false>(const void*, const void*, const void*, const void*, const void*, const void*, void*, void*, void*, void*, int64_t, int, int, RqsCfg<double>, cudaStream_t);
}
