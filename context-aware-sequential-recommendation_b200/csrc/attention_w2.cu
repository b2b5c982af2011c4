// explicit instantiations of the attention kernels for padded head widths 64, 4
#include "attention.cuh"

namespace cast {
template int dispatch_att<64>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
template int dispatch_att<4>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
}  // namespace cast
