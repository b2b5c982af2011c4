// explicit instantiations of the attention kernels for padded head widths 8, 12, 16
#include "attention.cuh"

namespace cast {
template int dispatch_att<8>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
template int dispatch_att<12>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
template int dispatch_att<16>(int, const AttnFwdArgs*, const AttnBwdArgs*, const AttnDims&, cudaStream_t);
}  // namespace cast
