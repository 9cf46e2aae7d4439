// temporary: tcgen05 family not built yet
#include "b2c_common.cuh"
namespace b2c {
bool tc_conv_supported(const ConvShape&, int) { return false; }
size_t tc_conv_workspace(const ConvShape&, int, int) { return 0; }
int launch_conv_tc(const ConvShape&, int, int, const float*, const float*, const float*, float*, void*, size_t, cudaStream_t) { return fail(B2C_ERR_INVALID, "tc not built"); }
bool tc_gemm_supported(bool, bool, int, int, int) { return false; }
int launch_sgemm_tc(bool, bool, int, int, int, float, const float*, const float*, float, float*, int, cudaStream_t) { return fail(B2C_ERR_INVALID, "tc not built"); }
}
