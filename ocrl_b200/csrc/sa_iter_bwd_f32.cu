// fp32 k/v instantiations of the fused iteration backward
#include "sa_iter_bwd.cuh"
namespace ocrl {
template int sa_iter_bwd_dispatch<float>(const IterBwdArgs&, cudaStream_t);
}
