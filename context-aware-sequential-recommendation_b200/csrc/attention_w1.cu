// explicit instantiations of the attention kernels for padded head widths 28, 32
#include "attention.cuh"

namespace cast {
template int dispatch_att<28>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
template int dispatch_att<32>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
}  // namespace cast
