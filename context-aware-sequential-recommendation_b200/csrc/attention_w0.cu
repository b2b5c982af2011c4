// explicit instantiations of the attention kernels for padded head widths 52
#include "attention.cuh"

namespace cast {
template int dispatch_att<52>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
}  // namespace cast
