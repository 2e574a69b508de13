// a7 register-path instantiations, float32
#include "transform_impl.cuh"
namespace nf {
template int rqs_unit_fwd_launch<float, false>(const void*, const void*, const void*, const void*, void*, void*, int64_t, int, int, RqsCfg<float>, cudaStream_t);
template int rqs_unit_bwd_launch<float, false>(const void*, const void*, const void*, const void*, const void*, const void*, void*, void*, void*, void*, int64_t, int, int, RqsCfg<float>, cudaStream_t);
}
