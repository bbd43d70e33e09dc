// bf16 k/v instantiations of the fused iteration backward
#include "sa_iter_bwd.cuh"
namespace ocrl {
template int sa_iter_bwd_dispatch<__nv_bfloat16>(const IterBwdArgs&, cudaStream_t);
}
