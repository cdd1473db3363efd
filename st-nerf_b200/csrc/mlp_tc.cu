// placeholder until the tcgen05 kernels land (next commit): every entry point fails loudly.
#include "mlp_tc.cuh"
namespace stnerf {
int tc_pack_spacenet(TcNet&, const float*, bool) { return STNERF_OK; }
int tc_pack_motionnet(TcNet&, const float*) { return STNERF_OK; }
void tc_free(TcNet&) {}
int tc_launch_spacenet(const PointSrc&, const TcNet&, const SpaceNetW&, int, float*, float*, float*, int, cudaStream_t) {
  return STNERF_EINVAL;
}
int tc_launch_motionnet(const PointSrc&, const TcNet&, const MotionNetW&, int, const int*, int, float*, float*, int,
                        cudaStream_t) {
  return STNERF_EINVAL;
}
}  // namespace stnerf
