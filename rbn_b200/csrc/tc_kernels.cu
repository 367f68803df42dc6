// placeholder until the tcgen05 path lands (next commit)
#include "common.cuh"
namespace rbn {
int tc_forward(const RbnDesc*, const WsLayout&, char*, const float*, const float*, const float*, const int*, const int*,
               float*, float*, float*, cudaStream_t) { set_error("tensor-core path not built"); return RBN_E_UNSUPPORTED; }
int tc_backward(const RbnDesc*, const WsLayout&, char*, const float*, const float*, const float*, const int*,
                const int*, const float*, float*, float*, float*, cudaStream_t) { set_error("tensor-core path not built"); return RBN_E_UNSUPPORTED; }
}
