// bf16 k/v instantiations of the fused iteration forward
#include "sa_iter_fwd.cuh"
namespace ocrl {
template int sa_iter_fwd_dispatch<__nv_bfloat16>(const IterFwdArgs&, cudaStream_t);
}
